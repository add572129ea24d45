"""GPU: the pass-boundary refit kernel against the reference's fitted GaussianMixture objects, and
the whole path through the drop-in predictor (`TraceWeaverV3.FindAssignments` signature) against
the reference's returned 6-tuple, on every golden fixture."""
import numpy as np
import pytest

from golden_util import Golden, golden_files

pytestmark = pytest.mark.gpu
FILES = golden_files(gpu=True)
IDS = [f.split("/")[-1][:-4] for f in FILES]

# nodejs terms (7-15 distinct delay values) whose BIC arg-min differs between scikit-learn and the
# device refit.  scikit-learn's diagonal covariance `avg(X^2) - mean^2 + 1e-6` cancels ~3.6e7 against
# itself there, and its OWN BIC moves by more than the K=4 / K=5 gap when the same samples are listed
# in another order (tests/gmm_conditioning.py prints the evidence; tests/test_oracle_gmm.py has the
# CPU-side list).  The device sums in warp order, so it lands on the other side for these terms.  The
# selection of these terms is not compared; the rest of the fixture is, with the reference's record
# substituted for the term so that everything downstream of the refit is still checked bit for bit.
ILL_CONDITIONED = {"node_load125__init-service": [0], "node_load125__service2": [1]}


class Span:
    """The four members of the reference's Span (spans.py:1-75) the path touches."""

    def __init__(self, trace_id, sid, start_mus, duration_mus):
        self.trace_id, self.sid, self.start_mus, self.duration_mus = trace_id, sid, int(start_mus), int(duration_mus)

    def GetId(self):
        return (self.trace_id, self.sid)


def reference_call_args(g: Golden):
    """Rebuild what executor.py:1172-1175 passed to FindAssignments for this fixture."""
    import networkx as nx
    z, m = g.z, g.meta
    in_spans = [Span(t, s, a, d) for t, s, a, d in zip(z["in_trace"], z["in_sid"], z["in_start"], z["in_dur"])]
    out_parts = {}
    for k, ep in enumerate(m["out_eps_given"]):
        out_parts[ep] = [Span(t, s, a, d) for t, s, a, d in
                         zip(z[f"out{k}_trace"], z[f"out{k}_sid"], z[f"out{k}_start"], z[f"out{k}_dur"])]
    G = nx.DiGraph()
    G.add_nodes_from(m["graph_nodes"])
    G.add_edges_from([tuple(e) for e in m["graph_edges"]])
    for ep in m["graph_nodes"]:
        assert [u for u, _ in G.in_edges(ep)] == m["graph_in_edges"][ep]
    truth = {}
    for e, ep in enumerate(g.topo):
        truth[ep] = {in_spans[i].GetId(): out_parts[ep][j].GetId() for i, j in enumerate(z["truth"][e]) if j >= 0}
    return {m["in_ep"]: in_spans}, out_parts, truth, G


@pytest.fixture(scope="module")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from traceweaver_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_refit_kernel_matches_sklearn(engine, path):
    import torch
    from traceweaver_b200 import refit
    from traceweaver_b200.batch import build_batch
    g = Golden(path)
    prob = g.problem()
    hb = build_batch([prob])
    engine.bind(hb)
    engine.prepare()
    n, E = prob.n_in, prob.E
    assign0 = np.full((E, n), -1, np.int32)
    mis0, idx0 = g.z["mis_rank"][0], g.z["topk_idx"][0]
    for i in range(n):
        if mis0[i] >= 0:
            assign0[:, i] = idx0[i, mis0[i]]
    dev = engine.device
    d, c = engine.delays(torch.from_numpy(assign0.reshape(-1)).to(dev))
    dt, ct = engine.delays(torch.from_numpy(np.ascontiguousarray(g.z["truth"]).reshape(-1)).to(dev))
    base = engine.gmm_stream_draws(dt, ct)
    given_pos = [g.topo.index(ep) for ep in g.meta["out_eps_given"]]
    order = torch.from_numpy(np.asarray(refit.reference_term_order(prob, given_pos), np.int32)).to(dev)
    prm, nsel = engine.gmm_refit(d, c, seed_select=g.meta["global_seed"], prob_base_skip=base, term_order=order,
                                 want_selected=True)
    engine.status()
    want = g.mix_table(prob)
    got = prm.table.cpu().numpy()
    keep = [t for t in range(len(want)) if t not in ILL_CONDITIONED.get(path.split("/")[-1][:-4], [])]
    assert np.array_equal(nsel.cpu().numpy()[keep], want[keep, 0].astype(np.int32))
    for t in keep:
        k = int(want[t, 0])
        np.testing.assert_allclose(got[t, 1:1 + k], want[t, 1:1 + k], rtol=1e-7)
        np.testing.assert_allclose(got[t, 6:6 + k], want[t, 6:6 + k], rtol=1e-7, atol=1e-7)  # mu*pc; mu may be 0
        np.testing.assert_allclose(got[t, 11:11 + k], want[t, 11:11 + k], atol=1e-7)
        np.testing.assert_allclose(got[t, 16:16 + k], want[t, 16:16 + k], atol=1e-7)


@pytest.fixture(scope="module")
def predictor():
    from traceweaver_b200.predictor import TraceWeaverV3
    return TraceWeaverV3({}, {}, device=0, seed_select=10)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_find_assignments_equals_reference(predictor, path):
    g = Golden(path)
    in_parts, out_parts, truth, G = reference_call_args(g)
    unstable = ILL_CONDITIONED.get(path.split("/")[-1][:-4], [])
    eng, orig_refit = predictor.engine, predictor.engine.gmm_refit
    if unstable:      # pin the reference's record of the ill-conditioned terms (see ILL_CONDITIONED)
        import torch
        from traceweaver_b200.batch import build_batch
        want_mix = g.mix_table(g.problem())

        def pinned_refit(*a, **k):
            prm = orig_refit(*a, **k)
            for t in unstable:
                prm.table[t].copy_(torch.from_numpy(want_mix[t]).to(prm.table.device))
            return prm
        eng.gmm_refit = pinned_refit
    try:
        res = predictor.FindAssignments("MaxScoreBatchSubsetWithSkips", g.meta["process"], in_parts, out_parts,
                                        False, [], truth, G)
    finally:
        if unstable:
            del eng.gmm_refit
    all_assign, all_topk, not_best, num_spans, per_span_cand, cnt_un = res
    z, m = g.z, g.meta
    in_ids = [s.GetId() for s in list(in_parts.values())[0]]
    n = len(in_ids)
    assert num_spans == m["num_spans"] == n
    assert cnt_un == m["cnt_unassigned"]
    assert not_best == m["not_best_count"]
    for e, ep in enumerate(g.topo):
        ids = [s.GetId() for s in out_parts[ep]]
        want = {in_ids[i]: (ids[j] if j >= 0 else ("NA", "NA")) for i, j in enumerate(z["assign"][e])}
        assert all_assign[ep] == want, ep                              # bit-exact parent indices
        want_topk = {in_ids[i]: [ids[z["topk_final"][i, r, e]] for r in range(z["topk_final_cnt"][i])]
                     for i in range(n)}
        assert all_topk[ep] == want_topk, ep
    assert [per_span_cand.get(i, 0) for i in in_ids] == z["per_span_candidates"].tolist()
    # log-likelihood scores of the final top-K lists within the north-star tolerance (1e-5 absolute);
    # the nodejs fixtures hold scores of magnitude 1e12 (a delay thousands of sigma from a narrow
    # component), whose f64 spacing alone is 5e-4: 1e-12 relative there
    got_s = predictor.last["topk_score"].cpu().numpy()
    np.testing.assert_allclose(got_s, z["topk2_score"][m["passes"] - 1], rtol=1e-12, atol=1e-5, equal_nan=True)


def test_unsupported_modes_fail_loudly(predictor):
    g = Golden(FILES[0])
    in_parts, out_parts, truth, G = reference_call_args(g)
    with pytest.raises(NotImplementedError):
        predictor.FindAssignments("MaxScoreBatchParallel", g.meta["process"], in_parts, out_parts, True, [], truth, G)
    with pytest.raises(NotImplementedError):   # the true-skips / true-distribution ablations (executor.py:1176-1183)
        predictor.FindAssignments("MaxScoreBatchSubsetWithSkips", g.meta["process"], in_parts, out_parts, False, [],
                                  truth, G, True, False)
    # a service with a skip budget (one outgoing span missing) is solved, not rejected (tests/test_skip_mode.py)
    ep = list(out_parts)[0]
    out_parts[ep] = out_parts[ep][:-1]
    res = predictor.FindAssignments("MaxScoreBatchSubsetWithSkips", g.meta["process"], in_parts, out_parts, False, [],
                                    truth, G)
    assert len(res) == 6 and res[3] == len(list(in_parts.values())[0])
