"""Host-side bookkeeping of the pass-boundary refit (ComputeEpPairDistParams5,
traceweaver_v3.py:706-818).

The fits themselves run on the device (tw_gmm_refit) — or, in `refit="sklearn"` parity mode, in
scikit-learn exactly as the reference calls it.  What lives here is the part that is pure
bookkeeping: the ORDER in which the reference fits the terms, which determines where in NumPy's
global random stream each model-selection fit starts (the reference never seeds that stream; the
golden harness and this package use np.random.seed(seed_select) right before FindAssignments).
"""
import math

import numpy as np

from . import _abi


def draws_for(max_n: int) -> int:
    """random_sample() calls consumed by the BIC fits n = 1..max_n of one term:
    k-means++ draws 1 + (n-1) * (2 + int(ln n)) per fit (sklearn/cluster/_kmeans.py)."""
    return sum(1 + (k - 1) * (2 + int(math.log(k))) for k in range(1, max_n + 1))


def reference_term_order(problem, given_pos):
    """Term indices (problem-local) in the order ComputeEpPairDistParams5 visits them:
    `for out_ep in out_span_partitions.keys()` (the GIVEN ep order, v3:801), per ep the root term,
    the primary in-edges, then the last term (v3:803-818) — i.e. the engine's per-ep term groups,
    re-ordered from topological to given ep order.  given_pos[g] = topological position of the
    g-th given ep."""
    terms = problem.terms()
    by_ep = {}
    for t, (e, _src) in enumerate(terms):
        by_ep.setdefault(e, []).append(t)
    order = []
    for e in given_pos:
        order.extend(by_ep[e])
    return order


def rng_skips(problem, given_pos, max_n_pred, max_n_truth=None):
    """Per-term random_sample() offset into the `seed_select` stream.

    The reference runs ComputeDistParams for the TRUE assignments first (v3:796, i = 0) and only
    then for the predicted ones; the truth fits do not influence results but consume
    draws_for(max_n_truth[t]) draws per term.  max_n_* = min(#unique delays, 5)
    per term (v3:768), problem-local term order.  max_n_truth=None: no truth pass (batch API)."""
    order = reference_term_order(problem, given_pos)
    pos = 0
    if max_n_truth is not None:
        for t in order:
            pos += draws_for(int(max_n_truth[t]))
    skips = np.zeros(len(order), np.uint32)
    for t in order:
        skips[t] = pos
        pos += draws_for(int(max_n_pred[t]))
    return skips


def unique_cap(delays, offsets, counts):
    """min(#unique, 5) per term (v3:768) from compacted delay samples."""
    out = np.zeros(len(counts), np.int32)
    for t, c in enumerate(counts):
        if c > 0:
            out[t] = min(len(np.unique(delays[offsets[t]:offsets[t] + c])), _abi.TW_GMM_MAX_COMP)
    return out
