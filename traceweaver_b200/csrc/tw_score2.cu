// tw_score2.cu — the scoring pass of tw_score_topk, work-balanced form.
//
// Same contract and results as k_score (tw_score.cu): FindTopKAssignments(K=5) on the undeleted
// lists (V3:1185: DfsTraverseX V3:292-351, ScoreAssignmentAsPerInvocationGraph V1:259-361,
// GetEpPairCost V1:117-139) plus the PerfectCut flags (V3:1024-1039).  k_score gives every
// in-span to one thread, and a depth-first walk whose length varies from 1 to >1000 tuples keeps
// ~2 of 32 lanes busy (profiles/ r01).  Here every phase after the per-in-span range search is
// indexed by WORK ITEM, not by in-span, so warps stay converged:
//   1. slots : every likelihood term value any tuple of the tile can need (term tables,
//              tw_core.cuh) is one slot; thread t evaluates slots t, t+T, ... (FP64 exp/log/div)
//   2. combos: every candidate combination of every in-span is one item; thread t decodes items
//              t, t+T, ..., tests feasibility, sums table entries, appends (score key, in-span,
//              combo) to a shared list, counts feasible tuples and marks candidate bitmaps
//   3. top-K : K rounds of a block-wide segmented arg-max over the list (64-bit atomicMax on the
//              order-preserving score key); winners write their rank straight to the outputs.
//              Exact score ties inside one in-span (rare) send that in-span to the sequential
//              walk so the reference's tie order is kept.
// In-spans whose tables or combination count exceed the shared-memory budgets go through the
// sequential walk too.  Windows-only launches and bitmap-overflow tiles stay on k_score.
#include "tw_kernels.cuh"

namespace tw {

// optional phase timers (build with -DTW_PROFILE_PHASES; scripts/phase_profile.py reads them)
#ifdef TW_PROFILE_PHASES
__device__ unsigned long long g_score2_phase[16];
#define TW_PHASE(k)                                                                  \
  do {                                                                               \
    if (threadIdx.x == 0) {                                                          \
      long long _now = clock64();                                                    \
      atomicAdd(&g_score2_phase[k], (unsigned long long)(_now - _phase_t0));         \
      _phase_t0 = _now;                                                              \
    }                                                                                \
  } while (0)
#else
#define TW_PHASE(k) do { } while (0)
#endif

// Shared-memory budgets of a tile.  They are sized for FOUR resident CTAs per SM (<= 55 KB each;
// the kernel needs 128 registers, which also allows four): the kernel is latency bound, and a
// third more warps bought more than the extra rounds of the heavier tiles cost.
#ifndef TW_S2_STAGE
#define TW_S2_STAGE 640
#define TW_S2_TBL 1152
#define TW_S2_ENT 576
#define TW_S2_PRM_TERMS 16
#endif
constexpr int kStage2 = TW_S2_STAGE;   // out spans staged per tile (else the tile reads global memory)
constexpr int kTbl2 = TW_S2_TBL;       // term-table slots per round
constexpr int kEnt2 = TW_S2_ENT;       // candidate combinations (= list entries) per round
constexpr int kPrm2 = TW_S2_PRM_TERMS * TW_MIX_REC;   // staged likelihood parameters (doubles)

template <int T>
struct Score2Smem {
  ProbView v;
  OutWin win[TW_MAX_E];
  int64_t st_s[kStage2];
  int64_t st_e[kStage2];
  double prm[kPrm2];
  double tbl[kTbl2];
  unsigned long long ent_key[kEnt2];
  unsigned long long rbest[T];
  int64_t ins[T], ine[T];
  int64_t red[T / 32];
  uint32_t ent_combo[kEnt2];
  uint32_t used[T][TW_MAX_E][kNarrowW];
  int lo_abs[T][TW_MAX_E];
  int rr[T][TW_MAX_E];
  int tstart[T + 1], cstart[T + 1];
  int rcount[T], nfeas[T];
  int scan_a[T / 32], scan_b[T / 32];
  int win_a[TW_MAX_E], win_n[TW_MAX_E];
  int staged, overflow, rc, n_ent, total_t, total_c, first_tid, last_tid, stream_tid, n_carry;
  unsigned long long carry_key[TW_K];
  uint32_t carry_combo[TW_K];
  uint16_t ent_j[kEnt2];
  uint8_t sid[kTbl2];
  uint8_t tie[T];
  double etab[64];
};

__device__ __forceinline__ double key_to_score(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)u);
}

#ifndef TW_S2_MINB
#define TW_S2_MINB 4
#endif
template <int T>
__global__ void __launch_bounds__(T, TW_S2_MINB)
k_score2(tw_batch b, tw_params prm, tw_score_out out, TileList tiles, const int32_t* __restrict__ prev_idx,
         uint8_t* __restrict__ overflow_flag, int* __restrict__ err_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Score2Smem<T>& sm = *reinterpret_cast<Score2Smem<T>*>(smem_raw);
  constexpr int W = kNarrowW;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#ifdef TW_PROFILE_PHASES
  long long _phase_t0 = clock64();
#endif
  const int t = blockIdx.x;
  const int p = tiles.tile_prob[t];
  const int i0 = tiles.tile_start[t];
  if (tid == 0) {
    sm.rc = load_view(b, p, sm.v);
    sm.overflow = 0;
    sm.n_ent = 0;
  }
  load_exp_table(sm.etab);
  TW_PHASE(0);
  if (sm.rc != TW_OK) {
    if (tid == 0) atomicMin(err_flag, sm.rc);
    return;
  }
  const ProbView& v = sm.v;
  const int n = v.n_in, E = v.E;
  const int cnt = min(tiles.tile_len, n - i0);
  const bool helper = (tid == T - 1) && (i0 >= 1);
  const bool worker = tid < cnt;
  int i = i0 + tid;
  if (helper) i = prev_idx[v.in_off + i0];
  int64_t in_s = 0, in_e = INT64_MIN;
  if (worker || helper) { in_s = v.is[i]; in_e = v.ie[i]; }
  sm.ins[tid] = in_s;
  sm.ine[tid] = in_e;
  sm.nfeas[tid] = 0;
  sm.tie[tid] = 0;
  sm.rbest[tid] = 0ULL;
  sm.rcount[tid] = 0;

  // ---- stage the tile's candidate slice of every ep (as k_score)
  int64_t me = worker ? in_e : INT64_MIN;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    int64_t o = __shfl_xor_sync(0xffffffffu, me, d);
    me = o > me ? o : me;
  }
  if (lane == 0) sm.red[wid] = me;
  __syncthreads();
  if (tid < E) {
    int64_t mx = sm.red[0];
    for (int q = 1; q < T / 32; ++q) mx = sm.red[q] > mx ? sm.red[q] : mx;
    int a = lower_bound(v.os[tid], v.n_out[tid], v.is[i0]);
    int z = upper_bound(v.os[tid], v.n_out[tid], mx);
    sm.win_a[tid] = a;
    sm.win_n[tid] = z > a ? z - a : 0;
  }
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int e = 0; e < E; ++e) tot += sm.win_n[e];
    sm.staged = tot <= kStage2;
    int off = 0;
    for (int e = 0; e < E; ++e) {
      if (sm.staged) {
        sm.win[e].s = sm.st_s + off; sm.win[e].e = sm.st_e + off;
        sm.win[e].base = sm.win_a[e]; sm.win[e].n = sm.win_n[e];
        off += sm.win_n[e];
      } else {
        sm.win[e].s = v.os[e]; sm.win[e].e = v.oe[e]; sm.win[e].base = 0; sm.win[e].n = v.n_out[e];
      }
    }
  }
  __syncthreads();
  if (sm.staged) {
    for (int e = 0; e < E; ++e) {
      const int64_t* gs = v.os[e] + sm.win_a[e];
      const int64_t* ge = v.oe[e] + sm.win_a[e];
      int64_t* ds = const_cast<int64_t*>(sm.win[e].s);
      int64_t* de = const_cast<int64_t*>(sm.win[e].e);
      for (int x = tid; x < sm.win_n[e]; x += T) { ds[x] = gs[x]; de[x] = ge[x]; }
    }
  }
  const int batch0 = i0 / TW_PARAM_BATCH;
  // likelihood parameters of the tile: staged when they fit, else read in place
  const double* prm_base;
  if (prm.mode == TW_PARAMS_GAUSS_BATCHED) {
    int nrec = 3 * v.n_terms * TW_GAUSS_REC;   // a 127-span tile can touch three 100-span batches
    const int nb = (n + TW_PARAM_BATCH - 1) / TW_PARAM_BATCH;
    const double* src = prm.gauss + (prm.prob_gauss_off[p] + (int64_t)batch0 * v.n_terms) * TW_GAUSS_REC;
    const int avail = (nb - batch0) * v.n_terms * TW_GAUSS_REC;
    if (nrec > avail) nrec = avail;
    prm_base = nrec <= kPrm2 ? sm.prm : src;
    if (nrec <= kPrm2)
      for (int x = tid; x < nrec; x += T) sm.prm[x] = src[x];
  } else {
    const double* src = prm.mix + (int64_t)v.term0 * TW_MIX_REC;
    const int nrec = v.n_terms * TW_MIX_REC;
    prm_base = nrec <= kPrm2 ? sm.prm : src;
    if (nrec <= kPrm2)
      for (int x = tid; x < nrec; x += T) sm.prm[x] = src[x];
  }
  for (int e = 0; e < TW_MAX_E; ++e)
    for (int wq = 0; wq < W; ++wq) sm.used[tid][e][wq] = 0u;
  __syncthreads();
  TW_PHASE(1);

  // ---- per in-span: candidate ranges, table size, number of combinations
  OutWin w[TW_MAX_E];
  int lo[TW_MAX_E], r[TW_MAX_E];
  int tsize = 0;
  long long P = 0;
  if (worker || helper) {
    for (int e = 0; e < E; ++e) {
      if (helper && sm.staged) { w[e].s = v.os[e]; w[e].e = v.oe[e]; w[e].base = 0; w[e].n = v.n_out[e]; }
      else w[e] = sm.win[e];
      lo[e] = lower_bound(w[e].s, w[e].n, in_s);
      sm.lo_abs[tid][e] = w[e].base + lo[e];
      r[e] = worker ? range_len(w[e], lo[e], in_e) : 0;
      sm.rr[tid][e] = r[e];
    }
    if (worker) { tsize = term_table_size(v, r); P = combo_count(v, r); }
  }
  const int* lo_abs = sm.lo_abs[tid];
  auto mark_serial = [&](const int* c, bool& ovf) {
    for (int e = 0; e < E; ++e) {
      int bit = c[e] - lo_abs[e];
      if (bit >= 32 * W) ovf = true;
      else atomicOr(&sm.used[tid][e][bit >> 5], 1u << (bit & 31));
    }
  };
  auto write_list = [&](const TopK& tk, int leaves) {
    const int64_t gi = v.in_off + i;
    out.n_feasible[gi] = leaves;
    out.topk_cnt[gi] = (uint8_t)tk.n;
    int32_t* ix = out.topk_idx + TW_K * (v.tuple_off + (int64_t)i * E);
    for (int k = 0; k < TW_K; ++k) {
      out.topk_score[gi * TW_K + k] = k < tk.n ? tk.score[k] : __longlong_as_double(0x7ff8000000000000LL);
      for (int e = 0; e < E; ++e) ix[k * E + e] = k < tk.n ? tk.idx[k][e] : -1;
    }
  };
  // sequential walk of one in-span by its owner thread (budget overflow, score ties)
  auto serial_walk = [&]() {
    ParamView pv;
    pv.mode = prm.mode;
    pv.gauss = prm_base + (i / TW_PARAM_BATCH - batch0) * v.n_terms * TW_GAUSS_REC;
    pv.mix = prm_base;
    pv.etab = sm.etab;
    TopK tk;
    tk.n = 0;
    int leaves = 0;
    bool ovf = false;
    enumerate(v, in_s, in_e, w, lo, [](int, int) { return false; },
              [&](const int* c, const int64_t* cs, const int64_t* ce) {
                if (leaves < 0x7fffffff) ++leaves;
                mark_serial(c, ovf);
                topk_offer(v, tk, score_tuple(v, pv, in_s, in_e, cs, ce), c);
              });
    if (ovf) sm.overflow = 1;
    write_list(tk, leaves);
  };
  if (helper) {   // the carry-in in-span only contributes its candidate bitmap
    bool ovf = false;
    enumerate(v, in_s, in_e, w, lo, [](int, int) { return false; },
              [&](const int* c, const int64_t*, const int64_t*) { mark_serial(c, ovf); });
    if (ovf) sm.overflow = 1;
  }

  TW_PHASE(2);
  bool pending = worker;
  while (true) {
    // ---- admit a prefix of the pending in-spans that fits the table and combination budgets
    const int my_t = pending ? tsize : 0;
    const int my_c = pending ? (int)(P > kEnt2 ? kEnt2 + 1 : P) : 0;
    int inc_t = my_t, inc_c = my_c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int ot = __shfl_up_sync(0xffffffffu, inc_t, d), oc = __shfl_up_sync(0xffffffffu, inc_c, d);
      if (lane >= d) { inc_t += ot; inc_c += oc; }
    }
    if (lane == 31) { sm.scan_a[wid] = inc_t; sm.scan_b[wid] = inc_c; }
    if (tid == 0) {
      sm.total_t = 0; sm.total_c = 0; sm.first_tid = T; sm.last_tid = -1; sm.n_ent = 0;
      sm.stream_tid = -1; sm.n_carry = 0;
    }
    __syncthreads();
    int toff = inc_t - my_t, coff = inc_c - my_c;
    for (int q = 0; q < wid; ++q) { toff += sm.scan_a[q]; coff += sm.scan_b[q]; }
    const bool fits = pending && toff + my_t <= kTbl2 && coff + my_c <= kEnt2;
    // The first pending in-span always makes progress.  More combinations than list entries: it
    // takes the round alone and its combinations are STREAMED through the list in chunks, the
    // running top-K carried from chunk to chunk.  Tables that cannot fit: sequential walk.
    const bool first_pending = pending && toff == 0 && coff == 0;
    const bool can_stream = tsize <= kTbl2 && P > kEnt2 && P < (1LL << 31);
    const bool stream = first_pending && can_stream;
    const bool serial = first_pending && !can_stream && (tsize > kTbl2 || P > kEnt2);
    const bool in_round = stream || (fits && !serial);
    sm.tstart[tid] = in_round ? toff : 0x3fffffff;
    sm.cstart[tid] = in_round ? coff : 0x3fffffff;
    if (in_round) {
      atomicMax(&sm.total_t, toff + my_t);
      atomicMax(&sm.total_c, stream ? (int)P : coff + my_c);
      atomicMin(&sm.first_tid, tid);
      atomicMax(&sm.last_tid, tid);
    }
    if (stream) sm.stream_tid = tid;
    if (serial) { serial_walk(); pending = false; }
    __syncthreads();
  TW_PHASE(3);
    const int total_t = sm.total_t, total_c = sm.total_c, f_tid = sm.first_tid, l_tid = sm.last_tid;
    const int stream_tid = sm.stream_tid;
    const int chunk = stream_tid >= 0 ? kEnt2 - TW_K : 0x7fffffff;
    auto owner_of = [&](const int* start, int item) {   // last in-round tid with start <= item
      int a = f_tid, z = l_tid + 1;
      while (a < z) {
        int mid = (a + z) >> 1;
        if (start[mid] <= item) a = mid + 1; else z = mid;
      }
      return a - 1;
    };
    // ---- 1. slots: all likelihood term values of the round
    for (int s = tid; s < total_t; s += T) {
      const int j = owner_of(sm.tstart, s);
      const int ij = i0 + j;
      int lo_rel[TW_MAX_E];
      for (int e = 0; e < E; ++e) lo_rel[e] = sm.lo_abs[j][e] - sm.win[e].base;
      ParamView pv;
      pv.mode = prm.mode;
      pv.gauss = prm_base + (ij / TW_PARAM_BATCH - batch0) * v.n_terms * TW_GAUSS_REC;
      pv.mix = prm_base;
      pv.etab = sm.etab;
      const int64_t je = sm.ine[j];
      double val = 0.0;
      const uint8_t id = term_slot_eval(v, pv, sm.ins[j], je, sm.win, lo_rel, sm.rr[j], s - sm.tstart[j],
                                        [&](int e, int x) { return sm.win[e].e[lo_rel[e] + x] <= je; }, &val);
      sm.sid[s] = id;
      sm.tbl[s] = val;
    }
    __syncthreads();
  TW_PHASE(4);
    // ---- 2 + 3 per chunk of combinations (one chunk unless an in-span is being streamed)
    for (long long cbase = 0; cbase == 0 || cbase < total_c; cbase += chunk) {
      const int cend = (int)(cbase + chunk < (long long)total_c ? cbase + chunk : (long long)total_c);
      if (cbase > 0) {   // streaming: the running top-K re-enters the list
        if (tid < sm.n_carry) {
          sm.ent_key[tid] = sm.carry_key[tid];
          sm.ent_combo[tid] = sm.carry_combo[tid];
          sm.ent_j[tid] = (uint16_t)stream_tid;
        }
        if (tid == 0) sm.n_ent = sm.n_carry;
        __syncthreads();
      }
      // ---- 2. combinations: feasibility, score, list entry, candidate bitmap
      for (int g = (int)cbase + tid; g < cend; g += T) {
        const int j = stream_tid >= 0 ? stream_tid : owner_of(sm.cstart, g);
        int lo_rel[TW_MAX_E], o_last[TW_MAX_E], c[TW_MAX_E];
        int64_t ce[TW_MAX_E];
        for (int e = 0; e < E; ++e) lo_rel[e] = sm.lo_abs[j][e] - sm.win[e].base;
        term_table_last_offsets(v, sm.rr[j], o_last);
        const int combo = g - sm.cstart[j];
        const double* tbl = sm.tbl + sm.tstart[j];
        if (!combo_feasible(v, sm.win, lo_rel, sm.rr[j], o_last, sm.sid + sm.tstart[j], combo, c, ce)) continue;
        atomicAdd(&sm.nfeas[j], 1);
        for (int e = 0; e < E; ++e) {
          int bit = c[e] - sm.lo_abs[j][e];
          if (bit >= 32 * W) sm.overflow = 1;
          else atomicOr(&sm.used[j][e][bit >> 5], 1u << (bit & 31));
        }
        unsigned long long key = score_key(table_score(v, sm.rr[j], sm.lo_abs[j], tbl, c, ce));
        if (key == 0ULL) key = 1ULL;
        const int slot = atomicAdd(&sm.n_ent, 1);
        sm.ent_key[slot] = key;
        sm.ent_combo[slot] = (uint32_t)combo;
        sm.ent_j[slot] = (uint16_t)j;
      }
      __syncthreads();
  TW_PHASE(5);
      // ---- 3. top-K: K rounds of segmented arg-max over the list
      const int n_ent = sm.n_ent;
      int ranks_done = 0;
      for (int rk = 0; rk < TW_K; ++rk) {
        bool any = false;
        for (int en = tid; en < n_ent; en += T) {
          const unsigned long long key = sm.ent_key[en];
          if (key != 0ULL) { atomicMax(&sm.rbest[sm.ent_j[en]], key); any = true; }
        }
        if (!__syncthreads_or(any)) break;
        ranks_done = rk + 1;
        for (int en = tid; en < n_ent; en += T) {
          const unsigned long long key = sm.ent_key[en];
          const int j = sm.ent_j[en];
          if (key != 0ULL && key == sm.rbest[j]) {
            if (atomicAdd(&sm.rcount[j], 1) == 0) {   // winner of rank rk for in-span j
              sm.ent_key[en] = 0ULL;
              const int ij = i0 + j;
              const int64_t gi = v.in_off + ij;
              out.topk_score[gi * TW_K + rk] = key_to_score(key);
              int32_t* ix = out.topk_idx + TW_K * (v.tuple_off + (int64_t)ij * E) + rk * E;
              unsigned idx = sm.ent_combo[en];
              for (int e = E - 1; e >= 0; --e) {
                const unsigned re = (unsigned)sm.rr[j][e], q = idx / re;
                ix[e] = sm.lo_abs[j][e] + (int)(idx - q * re);
                idx = q;
              }
              if (stream_tid >= 0) { sm.carry_key[rk] = key; sm.carry_combo[rk] = sm.ent_combo[en]; }
            } else {
              sm.tie[j] = 1;   // two tuples with the same score: keep the reference's tie order
            }
          }
        }
        __syncthreads();
        if (in_round) { sm.rbest[tid] = 0ULL; sm.rcount[tid] = 0; }
        __syncthreads();
      }
      if (tid == 0) sm.n_carry = ranks_done;
      __syncthreads();
  TW_PHASE(6);
      if (chunk == 0x7fffffff) break;
    }
    // ---- owners finish their in-span
    if (in_round) {
      const int nf = sm.nfeas[tid];
      if (sm.tie[tid]) {
        serial_walk();
      } else {
        const int64_t gi = v.in_off + i;
        const int kc = nf < TW_K ? nf : TW_K;
        out.n_feasible[gi] = nf;
        out.topk_cnt[gi] = (uint8_t)kc;
        int32_t* ix = out.topk_idx + TW_K * (v.tuple_off + (int64_t)i * E);
        for (int k = kc; k < TW_K; ++k) {
          out.topk_score[gi * TW_K + k] = __longlong_as_double(0x7ff8000000000000LL);
          for (int e = 0; e < E; ++e) ix[k * E + e] = -1;
        }
      }
      pending = false;
    }
    if (!__syncthreads_or(pending)) break;
  }
  __syncthreads();
  TW_PHASE(7);

  // ---- PerfectCut(i), V3:1034-1039, and the candidate maps for tw_stitch
  if (worker) {
    uint8_t cut = 0;
    if (i >= 1 && i <= n - 2) {
      const int pi = prev_idx[v.in_off + i];
      const int slot = pi >= i0 ? pi - i0 : T - 1;
      bool disjoint = true;
      for (int e = 0; e < E && disjoint; ++e)
        if (bitmaps_intersect(sm.used[slot][e], sm.lo_abs[slot][e], sm.used[tid][e], sm.lo_abs[tid][e], W))
          disjoint = false;
      cut = (uint8_t)(disjoint && v.ie[pi] <= in_e);
    }
    out.cut[v.in_off + i] = cut;
    if (out.used_lo) {
      const int64_t base = v.tuple_off + (int64_t)i * E;
      for (int e = 0; e < E; ++e) {
        out.used_lo[base + e] = sm.lo_abs[tid][e];
        out.used_bits[2 * (base + e)] = sm.used[tid][e][0];
        out.used_bits[2 * (base + e) + 1] = sm.used[tid][e][1];
      }
      out.used_wide[v.in_off + i] = 0;
    }
  }
  if (tid == 0 && sm.overflow) overflow_flag[t] = 1;
}

#ifdef TW_PROFILE_PHASES
extern "C" int tw_debug_score2_phases(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_score2_phase, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return -2;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_score2_phase, z, sizeof z);
  }
  return 0;
}
#endif

cudaError_t launch_score2(const tw_batch& b, const tw_params& prm, const tw_score_out& out, const TileList& narrow,
                          const int32_t* prev_idx, uint8_t* narrow_overflow, int* err_flag, cudaStream_t s) {
  using Sm = Score2Smem<kScoreThreads>;
  auto k = k_score2<kScoreThreads>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm));
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaError_t e = cudaMemsetAsync(narrow_overflow, 0, (size_t)narrow.n_tiles, s);
  if (e != cudaSuccess) return e;
  k<<<narrow.n_tiles, kScoreThreads, sizeof(Sm), s>>>(b, prm, out, narrow, prev_idx, narrow_overflow, err_flag);
  return cudaGetLastError();
}

}  // namespace tw
