/*
 * tw_oracle_driver.c — runs the restated path (tw_oracle.c + tw_oracle_gmm.c) over a whole batch,
 * one service per task, on a pool of host threads.  TEST INFRASTRUCTURE ONLY: this is the CPU
 * baseline bench.py reports (`cpu_baseline`, kind "port") and what `bench.py --impl reference`
 * times (the reference itself is Python + Gurobi and cannot travel to the GPU box), and the
 * checker smoke() compares the engine with.
 *
 * Per service = TraceWeaverV3.FindAssignments, traceweaver_v3.py:1087-1229:
 *   params0 (v3:580-646) -> cuts (v3:1020-1078) -> pass 0 (v3:1159-1219) -> delays + refit
 *   (v3:706-818) -> top-K on undeleted lists with the GMMs (v3:1185) -> pass 1.
 * The refit's model-selection stream starts at the seed for every service and visits the terms in
 * term order (the engine's convention when no ground truth / given ep order is supplied).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "../include/traceweaver_b200.h"
#include "tw_oracle.h"

int two_gmm_refit_ex(int32_t n_terms, const int64_t* term_sample_off, const double* delays,
                     const int32_t* counts, uint32_t seed_select, const uint32_t* rng_skip,
                     double* mix_out, int32_t* n_selected_out, int32_t* max_n_out);
int two_gmm_draws_for(int max_n);

typedef struct {
  const tw_batch* b;
  uint32_t seed;
  const tw_pass_out* final;
  const tw_score_out* topk_final;
  int32_t* n_cand_total;
  double* mix_out;
  const int64_t* prob_gauss_off;   /* [P+1] */
  const int64_t* term_sample_off;  /* [n_term_total+1] */
  double* gauss;                   /* scratch: [prob_gauss_off[P]][3] */
  double* delays;                  /* scratch: [term_sample_off[n_term_total]] */
  int32_t* counts;                 /* scratch: [n_term_total] */
  int32_t* assign0;                /* scratch: [prob_tuple_off[P]] */
  int8_t* mis0;                    /* scratch: [n_in_total] */
  int32_t* ncand0;                 /* scratch: [n_in_total] */
  uint8_t* cut;                    /* scratch: [n_in_total] */
  int32_t* nfeas;                  /* scratch: [n_in_total] */
  int32_t* counters0;              /* scratch: [P][4] */
  atomic_int next;
  atomic_int rc;
} job_t;

static int solve_one(job_t* j, int p) {
  const tw_batch* b = j->b;
  int rc;
  if ((rc = two_params_pass0(b, p, j->prob_gauss_off, j->gauss))) return rc;
  tw_params p0 = {TW_PARAMS_GAUSS_BATCHED, 0, j->prob_gauss_off, j->gauss, NULL};
  tw_score_out win = {NULL, NULL, NULL, j->nfeas, j->cut};
  if ((rc = two_score_problem(b, p, NULL, &win))) return rc;
  tw_pass_out r0 = {j->assign0, j->mis0, j->ncand0, NULL, NULL, NULL, j->counters0};
  if ((rc = two_stitch_problem(b, p, &p0, j->cut, &r0))) return rc;
  if ((rc = two_delays(b, p, j->assign0, j->term_sample_off, j->delays, j->counts))) return rc;
  int ep0 = b->prob_ep_off[p], ep1 = b->prob_ep_off[p + 1];
  int t0 = b->ep_term_off[ep0], t1 = b->ep_term_off[ep1];
  /* chain the model-selection stream over this service's terms */
  uint32_t skip[64];
  int32_t maxn[64];
  if (t1 - t0 > 64) return TW_ERR_INVALID;
  uint32_t pos = 0;
  for (int t = t0; t < t1; ++t) {
    /* max_n depends only on the samples: query it with a dry call on this single term */
    skip[t - t0] = pos;
    int32_t mx = 0, nsel = 0;
    uint32_t sk = pos;
    rc = two_gmm_refit_ex(1, j->term_sample_off + t, j->delays, j->counts + t, j->seed, &sk,
                          j->mix_out + (size_t)t * TW_MIX_REC, &nsel, &mx);
    if (rc) return rc;
    maxn[t - t0] = mx;
    pos += (uint32_t)two_gmm_draws_for(mx);
  }
  (void)skip; (void)maxn;
  tw_params p1 = {TW_PARAMS_MIXTURE, 0, j->prob_gauss_off, NULL, j->mix_out};
  if (j->topk_final && j->topk_final->topk_score) {
    tw_score_out top = *j->topk_final;
    if (!top.n_feasible) top.n_feasible = j->nfeas;
    if (!top.cut) top.cut = j->cut;
    if ((rc = two_score_problem(b, p, &p1, &top))) return rc;
  }
  if ((rc = two_stitch_problem(b, p, &p1, j->cut, j->final))) return rc;
  if (j->n_cand_total) {
    int64_t io = b->prob_in_off[p], n = b->prob_in_off[p + 1] - io;
    for (int64_t i = 0; i < n; ++i) j->n_cand_total[io + i] = j->ncand0[io + i] + j->final->n_cand[io + i];
  }
  return TW_OK;
}

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (;;) {
    int p = atomic_fetch_add(&j->next, 1);
    if (p >= j->b->n_problems) break;
    int rc = solve_one(j, p);
    if (rc) { int z = 0; atomic_compare_exchange_strong(&j->rc, &z, rc); }
  }
  return NULL;
}

int two_find_assignments(const tw_batch* b, uint32_t seed_select, int threads, const tw_pass_out* final,
                         const tw_score_out* topk_final, int32_t* n_cand_total, double* mix_out) {
  if (!b || !final || !final->assign || !final->mis_rank || !final->n_cand || !mix_out) return TW_ERR_INVALID;
  const int P = b->n_problems;
  job_t j;
  memset(&j, 0, sizeof j);
  j.b = b; j.seed = seed_select; j.final = final; j.topk_final = topk_final;
  j.n_cand_total = n_cand_total; j.mix_out = mix_out;
  int64_t* pgo = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P + 1));
  int64_t* tso = (int64_t*)malloc(sizeof(int64_t) * (size_t)(b->n_term_total + 1));
  pgo[0] = 0; tso[0] = 0;
  for (int p = 0; p < P; ++p) {
    int64_t n = b->prob_in_off[p + 1] - b->prob_in_off[p];
    int ep0 = b->prob_ep_off[p], ep1 = b->prob_ep_off[p + 1];
    int nt = b->ep_term_off[ep1] - b->ep_term_off[ep0];
    pgo[p + 1] = pgo[p] + ((n + TW_PARAM_BATCH - 1) / TW_PARAM_BATCH) * nt;
    for (int t = b->ep_term_off[ep0]; t < b->ep_term_off[ep1]; ++t) tso[t + 1] = tso[t] + n;
  }
  j.prob_gauss_off = pgo; j.term_sample_off = tso;
  j.gauss = (double*)malloc(sizeof(double) * (size_t)(pgo[P] * TW_GAUSS_REC + 1));
  j.delays = (double*)malloc(sizeof(double) * (size_t)(tso[b->n_term_total] + 1));
  j.counts = (int32_t*)malloc(sizeof(int32_t) * (size_t)(b->n_term_total + 1));
  j.assign0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(b->prob_tuple_off[P] + 1));
  j.mis0 = (int8_t*)malloc((size_t)b->n_in_total + 1);
  j.ncand0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(b->n_in_total + 1));
  j.cut = (uint8_t*)malloc((size_t)b->n_in_total + 1);
  j.nfeas = (int32_t*)malloc(sizeof(int32_t) * (size_t)(b->n_in_total + 1));
  j.counters0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)P * 4);
  atomic_init(&j.next, 0);
  atomic_init(&j.rc, 0);
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < threads - 1; ++t)
    if (pthread_create(&th[started], NULL, worker, &j) == 0) ++started;
  worker(&j);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(pgo); free(tso); free(j.gauss); free(j.delays); free(j.counts); free(j.assign0); free(j.mis0);
  free(j.ncand0); free(j.cut); free(j.nfeas); free(j.counters0);
  return atomic_load(&j.rc);
}
