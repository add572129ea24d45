// tw_emul.cpp — CPU stepping of the engine's per-thread device functions (tw_core.cuh).
//
// TEST INFRASTRUCTURE ONLY.  The build container has no GPU; this file compiles the very same
// __host__ __device__ functions the CUDA kernels call (enumerate, score_tuple, topk_offer,
// bitmaps_intersect, WindowCursor, mwis_solve) with g++ and walks the kernels' thread/warp loops
// sequentially, so logic errors are caught by `pytest -m "not gpu"` before a GPU call is spent.
// It is not linked into libtw_b200.so and the package never loads it: the product has no CPU
// path.  Orchestration mirrors tw_score.cu (k_score) and tw_stitch.cu (k_stitch).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../traceweaver_b200/csrc/tw_core.cuh"

using namespace tw;

// Slots available for term tables (shared memory in the kernels).  Tests set it small to force the
// lazy per-leaf path and large to force the table path.
static int g_table_cap = 4096;
extern "C" void twe_set_table_cap(int cap) { g_table_cap = cap; }
// In-spans with more than g_light_combos candidate combinations go through the lane-parallel path
// (32 lanes stride over the combos, partial top-K lists are merged); -1 disables it.
static long long g_light_combos = -1;
extern "C" void twe_set_light_combos(long long c) { g_light_combos = c; }

// one in-span through the table path: fill -> evaluate every slot -> DFS with look-ups
template <class Taken>
static long long enumerate_with_tables(const ProbView& v, const ParamView& pv, int64_t in_s, int64_t in_e,
                                       const OutWin* w, const int* lo, const int* r, int table_size, Taken taken,
                                       TopK& tk, uint32_t* mark, int W, int* overflow) {
  std::vector<double> tbl((size_t)table_size + 1);
  std::vector<uint8_t> sid((size_t)table_size + 1);
  int o_last[TW_MAX_E], lo_abs[TW_MAX_E];
  term_table_last_offsets(v, r, o_last);
  for (int e = 0; e < v.E; ++e) lo_abs[e] = w[e].base + lo[e];
  term_table_fill(v, in_s, in_e, w, lo, r, o_last, 0, taken, tbl.data(), sid.data());
  for (int s = 0; s < table_size; ++s)              // the dense, lane-parallel pass of the kernels
    if (sid[s] != TW_SLOT_INVALID) tbl[s] = term_logpdf(pv, sid[s] & 63, tbl[s]);
  long long leaves = 0;
  const long long P = combo_count(v, r);
  if (g_light_combos >= 0 && P > g_light_combos && P < (1LL << 40)) {
    const int L = 32;
    std::vector<TopK> part(L);
    for (int lane = 0; lane < L; ++lane) {
      part[lane].n = 0;
      enumerate_combos(v, w, lo, r, o_last, sid.data(), lane, L, P,
                       [&](const int* c, const int64_t* ce, long long) {
                         ++leaves;
                         if (mark)
                           for (int e = 0; e < v.E; ++e) {
                             int bit = c[e] - lo_abs[e];
                             if (bit >= 32 * W) { *overflow = 1; continue; }
                             mark[e * W + (bit >> 5)] |= 1u << (bit & 31);
                           }
                         topk_offer_sorted(v, part[lane], table_score(v, r, lo_abs, tbl.data(), c, ce), c);
                       });
    }
    int head[32] = {0};
    for (int k = 0; k < TW_K; ++k) {      // K rounds of "best head over the lanes"
      int best = -1;
      for (int lane = 0; lane < L; ++lane) {
        if (head[lane] >= part[lane].n) continue;
        if (best < 0 || cand_ahead(v, part[lane].score[head[lane]], part[lane].idx[head[lane]],
                                   part[best].score[head[best]], part[best].idx[head[best]]))
          best = lane;
      }
      if (best < 0) break;
      tk.score[tk.n] = part[best].score[head[best]];
      for (int e = 0; e < v.E; ++e) tk.idx[tk.n][e] = part[best].idx[head[best]][e];
      ++tk.n;
      ++head[best];
    }
    return leaves;
  }
  enumerate(v, in_s, in_e, w, lo,
            [&](int e, int o) { return sid[o_last[e] + (o - lo_abs[e])] == TW_SLOT_INVALID; },
            [&](const int* c, const int64_t*, const int64_t* ce) {
              ++leaves;
              if (mark)
                for (int e = 0; e < v.E; ++e) {
                  int bit = c[e] - lo_abs[e];
                  if (bit >= 32 * W) { *overflow = 1; continue; }
                  mark[e * W + (bit >> 5)] |= 1u << (bit & 31);
                }
              topk_offer(v, tk, table_score(v, r, lo_abs, tbl.data(), c, ce), c);
            });
  return leaves;
}

static ParamView param_view(const tw_params* prm, const ProbView& v, int p, int i) {
  ParamView pv;
  pv.mode = prm->mode;
  pv.gauss = nullptr;
  pv.mix = nullptr;
  if (prm->mode == TW_PARAMS_GAUSS_BATCHED)
    pv.gauss = prm->gauss + (prm->prob_gauss_off[p] + (int64_t)(i / TW_PARAM_BATCH) * v.n_terms) * TW_GAUSS_REC;
  else
    pv.mix = prm->mix + (int64_t)v.term0 * TW_MIX_REC;
  return pv;
}

extern "C" int twe_score_problem(const tw_batch* b, int p, const tw_params* prm, const tw_score_out* out,
                                 int W, int* overflow) {
  ProbView v;
  int rc = load_view(*b, p, v);
  if (rc) return rc;
  int n = v.n_in;
  // exclusive prefix arg-max of in_end, ties to the later index (k_prev_index)
  std::vector<int> prev(n, 0);
  for (int i = 1; i < n; ++i) {
    int pr = prev[i - 1];
    if (i >= 2 && v.ie[i - 1] >= v.ie[pr]) pr = i - 1;
    if (i == 1) pr = 0;
    prev[i] = pr;
  }
  std::vector<uint32_t> used((size_t)n * TW_MAX_E * W, 0u);
  std::vector<int> lo_abs((size_t)n * TW_MAX_E, 0);
  OutWin w[TW_MAX_E];
  for (int e = 0; e < v.E; ++e) w[e] = OutWin{v.os[e], v.oe[e], 0, v.n_out[e]};
  *overflow = 0;
  for (int i = 0; i < n; ++i) {
    int lo[TW_MAX_E];
    for (int e = 0; e < v.E; ++e) {
      lo[e] = lower_bound(w[e].s, w[e].n, v.is[i]);
      lo_abs[(size_t)i * TW_MAX_E + e] = lo[e];
    }
    TopK tk; tk.clear();
    long long leaves = 0;
    uint32_t* mine = &used[(size_t)i * TW_MAX_E * W];
    ParamView pv;
    if (prm) pv = param_view(prm, v, p, i);
    int64_t in_s = v.is[i], in_e = v.ie[i];
    int r[TW_MAX_E];
    for (int e = 0; e < v.E; ++e) r[e] = range_len(w[e], lo[e], in_e);
    int tsize = term_table_size(v, r);
    if (prm && tsize <= g_table_cap) {
      leaves = enumerate_with_tables(v, pv, in_s, in_e, w, lo, r, tsize, [](int, int) { return false; }, tk, mine,
                                     W, overflow);
    } else {
      enumerate(v, in_s, in_e, w, lo, [](int, int) { return false; },
                [&](const int* c, const int64_t* cs, const int64_t* ce) {
                  ++leaves;
                  for (int e = 0; e < v.E; ++e) {
                    int bit = c[e] - lo[e];
                    if (bit >= 32 * W) { *overflow = 1; continue; }
                    mine[e * W + (bit >> 5)] |= 1u << (bit & 31);
                  }
                  if (prm) topk_offer(v, tk, score_tuple(v, pv, in_s, in_e, cs, ce), c);
                });
    }
    topk_finish(v, tk);
    out->n_feasible[v.in_off + i] = (int32_t)leaves;
    if (prm && out->topk_score) {
      out->topk_cnt[v.in_off + i] = (uint8_t)tk.n;
      for (int k = 0; k < TW_K; ++k) {
        out->topk_score[(v.in_off + i) * TW_K + k] = k < tk.n ? tk.score[k] : NAN;
        for (int e = 0; e < v.E; ++e)
          out->topk_idx[TW_K * (v.tuple_off + (int64_t)i * v.E) + k * v.E + e] = k < tk.n ? tk.idx[k][e] : -1;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    uint8_t cut = 0;
    if (i >= 1 && i <= n - 2) {
      int pi = prev[i];
      bool disjoint = true;
      for (int e = 0; e < v.E && disjoint; ++e)
        if (bitmaps_intersect(&used[((size_t)pi * TW_MAX_E + e) * W], lo_abs[(size_t)pi * TW_MAX_E + e],
                              &used[((size_t)i * TW_MAX_E + e) * W], lo_abs[(size_t)i * TW_MAX_E + e], W))
          disjoint = false;
      cut = (uint8_t)(disjoint && v.ie[pi] <= v.ie[i]);
    }
    out->cut[v.in_off + i] = cut;
  }
  return TW_OK;
}

extern "C" int twe_stitch_problem(const tw_batch* b, int p, const tw_params* prm, const uint8_t* cut_all,
                                  const tw_pass_out* out, long long node_limit) {
  ProbView v;
  int rc = load_view(*b, p, v);
  if (rc) return rc;
  int n = v.n_in;
  const uint8_t* cut = cut_all + v.in_off;
  std::vector<std::vector<uint8_t>> taken(v.E);
  for (int e = 0; e < v.E; ++e) taken[e].assign((size_t)v.n_out[e], 0);
  OutWin w[TW_MAX_E];
  for (int e = 0; e < v.E; ++e) w[e] = OutWin{v.os[e], v.oe[e], 0, v.n_out[e]};
  for (int e = 0; e < v.E; ++e)
    for (int i = 0; i < n; ++i) out->assign[v.tuple_off + (int64_t)e * n + i] = -1;
  for (int i = 0; i < n; ++i) out->mis_rank[v.in_off + i] = -1;
  WindowCursor wc; wc.init();
  WindowBuf* wb = new WindowBuf;
  int cursor[TW_MAX_E] = {0};
  int not_best = 0, unassigned = 0; long long max_nodes = 0;
  int ws = 0;
  while (ws < n) {
    int we = ws;
    while (!wc.ends_at(we, n, cut) && we < n - 1) ++we;
    bool closes = (we != n - 1) || (n > 1);
    int nw = we - ws + 1;
    if (nw > TW_WINDOW_CAP) { delete wb; return TW_ERR_INVALID; }
    for (int l = 0; l < nw; ++l) {  // one lane per in-span of the window
      int i = ws + l;
      int lo[TW_MAX_E];
      for (int e = 0; e < v.E; ++e) lo[e] = lower_bound_from(w[e].s, w[e].n, cursor[e], v.is[i]);
      if (l == 0) for (int e = 0; e < v.E; ++e) cursor[e] = lo[e];
      TopK tk; tk.clear();
      long long leaves = 0;
      ParamView pv = param_view(prm, v, p, i);
      int64_t in_s = v.is[i], in_e = v.ie[i];
      int r[TW_MAX_E];
      for (int e = 0; e < v.E; ++e) r[e] = range_len(w[e], lo[e], in_e);
      int tsize = term_table_size(v, r);
      auto is_taken = [&](int e, int o) { return taken[e][o] != 0; };
      if (tsize <= g_table_cap) {
        int ov = 0;
        leaves = enumerate_with_tables(v, pv, in_s, in_e, w, lo, r, tsize, is_taken, tk, nullptr, 0, &ov);
      } else {
        enumerate(v, in_s, in_e, w, lo, is_taken,
                  [&](const int* c, const int64_t* cs, const int64_t* ce) {
                    ++leaves;
                    topk_offer(v, tk, score_tuple(v, pv, in_s, in_e, cs, ce), c);
                  });
      }
      topk_finish(v, tk);
      out->n_cand[v.in_off + i] = (int32_t)leaves;
      wb->cnt[l] = tk.n;
      for (int k = 0; k < tk.n; ++k) {
        wb->score[l][k] = tk.score[k];
        for (int e = 0; e < v.E; ++e) wb->idx[l][k][e] = tk.idx[k][e];
      }
      if (out->topk_score) {
        out->topk_cnt[v.in_off + i] = (uint8_t)tk.n;
        for (int k = 0; k < TW_K; ++k) {
          out->topk_score[(v.in_off + i) * TW_K + k] = k < tk.n ? tk.score[k] : NAN;
          for (int e = 0; e < v.E; ++e)
            out->topk_idx[TW_K * (v.tuple_off + (int64_t)i * v.E) + k * v.E + e] = k < tk.n ? tk.idx[k][e] : -1;
        }
      }
    }
    if (closes) {
      for (int l = 0; l < nw; ++l) wb->adj[l] = window_adjacency(*wb, v.E, nw, l);
      long long nodes = mwis_solve(*wb, v.E, nw, node_limit);
      if (nodes < 0) { delete wb; return TW_ERR_MWIS_LIMIT; }
      if (nodes > max_nodes) max_nodes = nodes;
      for (int l = 0; l < nw; ++l) {
        int i = ws + l, r = wb->chosen[l];
        out->mis_rank[v.in_off + i] = (int8_t)r;
        if (r != 0) ++not_best;
        if (r < 0) { ++unassigned; continue; }
        for (int e = 0; e < v.E; ++e) {
          int o = wb->idx[l][r][e];
          out->assign[v.tuple_off + (int64_t)e * n + i] = o;
          taken[e][o] = 1;
        }
      }
    }
    ws = we + 1;
  }
  if (out->counters) {
    out->counters[p * 4 + 0] = not_best;
    out->counters[p * 4 + 1] = unassigned;
    out->counters[p * 4 + 2] = (int32_t)(max_nodes > 0x7fffffff ? 0x7fffffff : max_nodes);
    out->counters[p * 4 + 3] = 0;
  }
  delete wb;
  return TW_OK;
}

// exact matching of one E = 1 window given as candidate lists (unit test against scipy)
extern "C" int twe_assign_window(int nw, const int* cnt, const double* score, const int* span, int* chosen) {
  WindowBuf* wb = new WindowBuf;
  int member[TW_WINDOW_CAP];
  for (int k = 0; k < nw; ++k) {
    wb->cnt[k] = cnt[k];
    member[k] = k;
    for (int r = 0; r < cnt[k]; ++r) { wb->score[k][r] = score[k * TW_K + r]; wb->idx[k][r][0] = span[k * TW_K + r]; }
  }
  assignment_solve(*wb, member, nw, 0, chosen, nullptr, nullptr);
  delete wb;
  return 0;
}

// one window through the sequential solver of the stitch kernel (adjacency + components + branch and
// bound / Hungarian); returns the node count (-1 = node limit)
extern "C" long long twe_mwis_window(int nw, int E, const int* cnt, const double* score, const int* idx,
                                     int* chosen, long long node_limit) {
  WindowBuf* wb = new WindowBuf;
  for (int k = 0; k < nw; ++k) {
    wb->cnt[k] = cnt[k];
    for (int r = 0; r < cnt[k]; ++r) {
      wb->score[k][r] = score[k * TW_K + r];
      for (int e = 0; e < E; ++e) wb->idx[k][r][e] = idx[(k * TW_K + r) * E + e];
    }
  }
  for (int k = 0; k < nw; ++k) wb->adj[k] = window_adjacency(*wb, E, nw, k);
  long long nodes = mwis_solve(*wb, E, nw, node_limit);
  for (int k = 0; k < nw; ++k) chosen[k] = wb->chosen[k];
  delete wb;
  return nodes;
}

// The warp-parallel solver of tw_stitch.cu (stitch_small_window) restated for one thread: candidate
// (k, r) = "lane" 5k + r, conflict masks per candidate, components of the in-span conflict graph,
// exhaustive mixed-radix enumeration (lowest in-span most significant, digit cnt = unassigned),
// the lowest combination index among the totals tied with the largest (TW_MWIS_TIE_TOL).  Returns 0 when the kernel would fall back
// to the sequential solver (E == 1 component of >= 3 in-spans, or more than `space_cap` leaves).
extern "C" int twe_small_window(int nw, int E, const int* cnt_in, const double* score, const int* idx,
                                int* chosen, int space_cap) {
  const int KW = 6;
  if (nw > KW) return 0;
  uint32_t conf[32] = {0};
  double cw[32] = {0};
  auto valid = [&](int c) { int k = c / TW_K, r = c % TW_K; return k < nw && r < cnt_in[k]; };
  for (int a = 0; a < 32; ++a) {
    if (!valid(a)) continue;
    cw[a] = TW_WEIGHT_OFFSET + score[a];
    for (int b = 0; b < 32; ++b) {
      if (!valid(b) || a / TW_K == b / TW_K) continue;
      for (int e = 0; e < E; ++e)
        if (idx[a * E + e] == idx[b * E + e]) conf[a] |= 1u << b;
    }
  }
  uint32_t adj[KW] = {0};
  for (int a = 0; a < 32; ++a)
    for (int q = 0; q < KW; ++q)
      if ((conf[a] >> (TW_K * q)) & 0x1fu) adj[a / TW_K] |= 1u << q;
  for (int k = 0; k < nw; ++k) chosen[k] = -1;
  uint32_t todo = (1u << nw) - 1u;
  while (todo) {
    int seed = 0;
    while (!(todo >> seed & 1u)) ++seed;
    uint32_t comp = 1u << seed, frontier = comp;
    while (frontier) {
      int q = 0;
      while (!(frontier >> q & 1u)) ++q;
      frontier &= frontier - 1u;
      uint32_t nb = adj[q] & ~comp;
      comp |= nb;
      frontier |= nb;
    }
    todo &= ~comp;
    int m = 0;
    for (int a = 0; a < KW; ++a) m += (comp >> a) & 1u;
    if (m == 1) {
      if (cnt_in[seed] > 0 && cw[TW_K * seed] > 0.0) chosen[seed] = 0;
      continue;
    }
    if (E == 1 && m >= 3) return 0;
    int stride[KW], space = 1;
    for (int a = KW - 1; a >= 0; --a) {
      stride[a] = space;
      if ((comp >> a) & 1u) space *= (a < nw ? cnt_in[a] : 0) + 1;
    }
    if (space > space_cap) return 0;
    // two passes like the kernel: the maximum total, then the lowest leaf index tied with it
    auto leaf_total = [&](int id, bool& ok) {
      int rem = id;
      uint32_t sel = 0;
      double tot = 0.0;
      ok = true;
      for (int a = 0; a < KW; ++a) {
        if (!((comp >> a) & 1u)) continue;
        int d = rem / stride[a];
        rem -= d * stride[a];
        if (d < cnt_in[a]) {
          int c = TW_K * a + d;
          if (!(cw[c] > 0.0) || (conf[c] & sel)) ok = false;
          sel |= 1u << c;
          tot = tot + cw[c];
        }
      }
      return tot;
    };
    double best_w = -1.0;
    for (int id = 0; id < space; ++id) {
      bool ok;
      const double tot = leaf_total(id, ok);
      if (ok && tot > best_w) best_w = tot;
    }
    int best_idx = 0x7fffffff;
    for (int id = 0; id < space && best_idx == 0x7fffffff; ++id) {
      bool ok;
      const double tot = leaf_total(id, ok);
      if (ok && tot >= best_w - TW_MWIS_TIE_TOL) best_idx = id;
    }
    int rem = best_idx;
    for (int a = 0; a < KW; ++a) {
      if (!((comp >> a) & 1u)) continue;
      int d = rem / stride[a];
      rem -= d * stride[a];
      chosen[a] = d < cnt_in[a] ? d : -1;
    }
  }
  return 1;
}

// ---------------------------------------------------------------------------------------------
// skip / cache mode (tw_skip_core.cuh): the kernel body of k_skip and k_build_dist on the CPU
// ---------------------------------------------------------------------------------------------
#include "../../traceweaver_b200/csrc/tw_skip_core.cuh"

extern "C" int twe_skip_solve(const tw_batch* b, const tw_skip_desc* sd, const tw_skip_out* out, long long node_limit) {
  const int P = b->n_problems;
  std::vector<int64_t> set_off((size_t)P + 1, 0);
  for (int p = 0; p < P; ++p) {
    int64_t words = 0;
    for (int ep = b->prob_ep_off[p]; ep < b->prob_ep_off[p + 1]; ++ep)
      words += (b->ep_out_off[ep + 1] - b->ep_out_off[ep] + 31) / 32;
    set_off[p + 1] = set_off[p] + words;
  }
  std::vector<uint32_t> sets((size_t)(3 * set_off[P]) + 1), taken((size_t)(b->n_out_total / 32) + (size_t)b->n_ep_total + 2);
  std::vector<int32_t> win((size_t)(2 * sd->prob_cnt_off[P]) + 1);
  int rc = TW_OK;
  SkipShared* sh = new SkipShared;
  for (int p = 0; p < P; ++p) {
    const int r = skip_solve_problem(*b, p, *sd, *out, taken.data(), sets.data(), set_off.data(), win.data(), node_limit, *sh);
    if (r < rc) rc = r;
  }
  delete sh;
  return rc;
}
