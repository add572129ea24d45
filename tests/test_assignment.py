"""E = 1 windows are maximum-weight bipartite matchings: the engine's Hungarian solver
(tw_core.cuh: assignment_solve, compiled for the CPU by tests/emul) against
scipy.optimize.linear_sum_assignment on dense random windows of up to 31 in-spans — the regime
where branch and bound over (in-span, rank) vertices explodes."""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

import emul_backend


def _solve(cnt, score, span):
    lib = emul_backend.lib()
    nw = len(cnt)
    chosen = np.full(nw, -9, np.int32)
    lib.twe_assign_window(nw, np.ascontiguousarray(cnt, np.int32).ctypes.data_as(C.c_void_p),
                          np.ascontiguousarray(score, np.float64).ctypes.data_as(C.c_void_p),
                          np.ascontiguousarray(span, np.int32).ctypes.data_as(C.c_void_p),
                          chosen.ctypes.data_as(C.c_void_p))
    return chosen


@pytest.mark.parametrize("seed", range(40))
def test_hungarian_matches_scipy(seed):
    rng = np.random.default_rng(seed)
    nw = int(rng.integers(3, 32))
    n_spans = int(rng.integers(max(2, nw // 2), nw + 6))          # scarce spans: heavy competition
    cnt = np.minimum(rng.integers(0, 6, nw), n_spans)
    score = np.full((nw, 5), np.nan)
    span = np.full((nw, 5), -1, np.int32)
    for k in range(nw):
        span[k, :cnt[k]] = rng.choice(n_spans, cnt[k], replace=False)
        score[k, :cnt[k]] = -np.sort(rng.exponential(30.0, cnt[k]))   # descending, like a top-K list
    if seed % 7 == 0 and cnt[0] > 0:
        score[0, 0] = -20000.0                                      # vertex weight <= 0 is never taken
    chosen = _solve(cnt, score, span)
    # reference optimum: rows = in-spans, columns = spans + one private "unassigned" column each
    W = np.zeros((nw, n_spans + nw))
    for k in range(nw):
        for r in range(cnt[k]):
            w = 10000.0 + score[k, r]
            if w > 0:
                W[k, span[k, r]] = max(W[k, span[k, r]], w)
    rows, cols = linear_sum_assignment(-W)
    best = W[rows, cols].sum()
    got, used = 0.0, set()
    for k in range(nw):
        r = chosen[k]
        assert -1 <= r < max(cnt[k], 1)
        if r >= 0:
            assert 10000.0 + score[k, r] > 0
            assert span[k, r] not in used
            used.add(span[k, r])
            got += 10000.0 + score[k, r]
    assert got == pytest.approx(best, abs=1e-6)
