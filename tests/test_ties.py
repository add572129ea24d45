"""Exact (score, start) ties: the reference keeps its top-K in a heapq of (score, stack) tuples
(traceweaver_v3.py:305-307) and sorts it with list.sort(reverse=True) (:461).  Two different spans
with the same start_mus compare neither-smaller (spans.py:51), so which of two equal entries
survives and where it lands is decided by heapq's sift order.  The oracle replays heapq; the
engine's sequential paths (tw_core.cuh topk_offer / topk_finish, stepped here on the CPU) must
give the same lists.  Inputs: millisecond clocks (alibaba-shaped), parallel eps, heavy overlap."""
import numpy as np
import pytest

from oracle import tw_oracle
from emul_backend import EmulBatch
from traceweaver_b200 import synth
from traceweaver_b200.batch import build_batch_from_blocks

CASES = {
    "par3_ms": ("ali_par3", 3, 160, 100.0, 1000),
    "chain2_ms": ("ali_chain2", 3, 200, 100.0, 1000),
    "leaf_ms": ("ali_leaf", 3, 200, 100.0, 1000),
    "leaf_dense_ms": ("ali_leaf", 4, 300, 300.0, 1000),      # E = 1: tied optimal matchings (Hungarian path)
    "nginx_2ms": ("media_nginx_cal", 2, 120, 60.0, 2000),
}


def _batch(name):
    shape, S, n, load, q = CASES[name]
    return build_batch_from_blocks([synth.make_block(shape, S, n, load, seed=77, quantum_us=q)])


@pytest.mark.parametrize("name", list(CASES))
def test_tie_order_is_the_references(name):
    hb = _batch(name)
    ob, eb = tw_oracle.OracleBatch(hb), EmulBatch(hb)
    g = ob.params_pass0()
    o_sc, e_sc = ob.score(gauss=g), eb.score(gauss=g)
    # the input really has ties inside the kept lists: equal scores at neighbouring ranks
    s = o_sc["topk_score"]
    assert int((s[:, 1:] == s[:, :-1]).sum()) > 0, "no exact score ties in this input"
    for k in ("cut", "n_feasible", "topk_cnt", "topk_idx"):
        assert np.array_equal(e_sc[k], o_sc[k]), k
    assert np.array_equal(np.nan_to_num(e_sc["topk_score"]), np.nan_to_num(o_sc["topk_score"]))
    o_st, e_st = ob.stitch(o_sc["cut"], gauss=g), eb.stitch(o_sc["cut"], gauss=g)
    for k in ("assign", "mis_rank", "n_cand", "topk_idx", "topk_cnt"):
        assert np.array_equal(e_st[k], o_st[k]), k
    assert np.array_equal(e_st["counters"][:, :2], o_st["counters"][:, :2])
