"""GPU: ground truth, FindOrder and the accuracies on the device (row f-2, csrc/tw_truth.cu) against the
fixtures minted from the reference (truth, invocation graph) and against the NumPy helpers of
traceweaver_b200.loader, which tests/test_accuracy.py pins to the accuracies the reference printed."""
import glob
import json
import os

import numpy as np
import pytest

from golden_util import Golden, GOLDEN_DIR

pytestmark = pytest.mark.gpu
DATASETS = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*_load*.json")))
IDS = [os.path.basename(p)[:-5] for p in DATASETS]


@pytest.fixture(scope="module")
def engine():
    from traceweaver_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def _dataset(path):
    meta = json.load(open(path))
    order = list(meta["find_assignments_seconds"].keys())              # the order the services were solved in
    return meta, [Golden(os.path.join(GOLDEN_DIR, f"{meta['dataset']}__{n}.npz")) for n in order]


def _trace_lists(gs):
    """Lists in the order the reference's loader produced them (callees as GIVEN, before FindOrder)."""
    from traceweaver_b200.truth import TraceLists
    number = {}
    probs, in_tr, out_tr = [], [], []
    for g in gs:
        z = g.z
        E = g.E
        in_tr.append(np.array([number.setdefault(str(t), len(number)) for t in z["in_trace"]], np.int32))
        out_tr.append([np.array([number.setdefault(str(t), len(number)) for t in z[f"out{k}_trace"]], np.int32)
                       for k in range(E)])
        ins = z["in_start"].astype(np.int64)
        probs.append(dict(in_start=ins, in_end=ins + z["in_dur"].astype(np.int64),
                          out_start=[z[f"out{k}_start"].astype(np.int64) for k in range(E)],
                          out_end=[(z[f"out{k}_start"] + z[f"out{k}_dur"]).astype(np.int64) for k in range(E)]))
    return TraceLists(probs, in_tr, out_tr, len(number)), number


@pytest.mark.parametrize("path", DATASETS, ids=IDS)
def test_truth_and_order_equal_reference(engine, path):
    from traceweaver_b200 import truth as T
    meta, gs = _dataset(path)
    tl, _ = _trace_lists(gs)
    truth = T.ground_truth(engine, tl)
    viol = T.find_order(engine, tl, truth)
    truth = truth.cpu().numpy()
    for p, g in enumerate(gs):
        E, n = g.E, len(g.z["in_start"])
        t0, e0 = int(tl.arrays["prob_tuple_off"][p]), int(tl.arrays["prob_ep_off"][p])
        got = truth[t0:t0 + E * n].reshape(E, n)                       # given callee order
        want = np.empty_like(got)
        for e_topo, k_given in enumerate(g.pos_given):
            want[k_given] = g.z["truth"][e_topo]
        assert np.array_equal(got, want), g.name
        given = g.meta["out_eps_given"]
        kept = {(given[a], given[b]) for a in range(E) for b in range(E) if a != b and not (int(viol[e0 + a]) >> b & 1)}
        assert kept == {tuple(e) for e in g.meta["graph_edges"]}, g.name


@pytest.mark.parametrize("path", DATASETS, ids=IDS)
def test_accuracies_equal_reference_helpers(engine, path):
    import torch
    from traceweaver_b200 import truth as T
    from traceweaver_b200.batch import build_batch
    from traceweaver_b200.loader import topk_accuracy, end_to_end_accuracy, end_to_end_topk_accuracy
    meta, gs = _dataset(path)
    hb = build_batch([g.problem() for g in gs])
    number = {}
    in_trace = np.concatenate([[number.setdefault(str(t), len(number)) for t in g.z["in_trace"]] for g in gs]).astype(np.int32)
    tl = T.TraceLists.from_host_batch(hb, in_trace, len(number))
    dev = engine.device
    truth = torch.from_numpy(np.concatenate([g.z["truth"].reshape(-1) for g in gs]).astype(np.int32)).to(dev)
    assign = torch.from_numpy(np.concatenate([g.z["assign"].reshape(-1) for g in gs]).astype(np.int32)).to(dev)
    tki = torch.from_numpy(np.concatenate([g.z["topk_final"].reshape(-1) for g in gs]).astype(np.int32)).to(dev)
    tkc = torch.from_numpy(np.concatenate([g.z["topk_final_cnt"] for g in gs]).astype(np.uint8)).to(dev)
    first = np.zeros(len(gs), np.uint8)
    first[0] = 1
    res = T.accuracy(engine, tl, truth, assign, tki, tkc, prob_first=first)
    for p, g in enumerate(gs):
        z = g.z
        ok = (z["assign"] == z["truth"]).all(axis=0)
        assert res["correct"][p] == int(ok.sum())
        assert abs(res["service_topk_accuracy"][p] - topk_accuracy(z["truth"], z["topk_final"], z["topk_final_cnt"])) < 1e-12
    tr = [list(g.z["in_trace"]) for g in gs]
    tru = [g.z["truth"] for g in gs]
    assert abs(res["e2e_accuracy"] - end_to_end_accuracy(tr, tru, [g.z["assign"] for g in gs])) < 1e-12
    want_k = end_to_end_topk_accuracy(tr, tru, [g.z["topk_final"] for g in gs], [g.z["topk_final_cnt"] for g in gs])
    assert abs(res["e2e_topk_accuracy"] - want_k) < 1e-12
