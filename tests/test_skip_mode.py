"""Skip / cache mode.

CPU: the literal oracle (oracle/tw_oracle_skip.py) and the host mirror's NumPy pieces
(traceweaver_b200/skipmode.py) against fixtures minted from the reference run with --cache_rate
(tests/golden_cache/, tests/golden/make_goldens.py <dataset>@<rate>).
GPU: tw_skip_solve / tw_build_dist_samples through the C ABI against the same fixtures and the oracle."""
import glob
import os

import numpy as np
import pytest

from golden_util import Golden
from oracle import tw_oracle_skip as osk
from traceweaver_b200 import skipmode

CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cache")


def cache_files(skip_only=True):
    out = []
    for f in sorted(glob.glob(os.path.join(CACHE_DIR, "*__*.npz"))):
        g = Golden(f)
        has_skip = any(v != 0 for v in g.meta["skip_budget"].values())
        if has_skip or not skip_only:
            out.append(f)
    return out


FILES = cache_files()
IDS = [os.path.basename(f)[:-4] for f in FILES]


def _inputs(g):
    prob = g.problem()          # lists in the caller's order (the cache transform leaves them partly unsorted)
    wins_before = [tuple(w) for w in g.meta["time_windows_before"]]
    return prob, wins_before


def _check_against_golden(g, res, exact_scores):
    z, m = g.z, g.meta
    assert [tuple(w) for w in m["time_windows"]] == [tuple(w) for w in res["time_windows"]]
    assert [m["skip_budget"][ep] for ep in g.topo] == list(res["skip_budget"])
    assert np.array_equal(np.asarray([m["skip_count"][ep] for ep in g.topo]), np.asarray(res["skip_count"]))
    labels = [m["in_ep"]] + g.topo
    for k, (mu, sd) in m["build_dist"].items():
        a, b = k.split("|")
        if a in labels and b in labels:
            got = res["pair_params"][labels.index(a), labels.index(b)]
            assert got[0] == mu and got[1] == sd, k
    assert res["large_delay"] == m["large_delay"]
    assert np.array_equal(res["topk2_idx"], z["topk2_idx"][0])
    assert np.array_equal(res["topk_idx"], z["topk_idx"][0])
    assert np.array_equal(res["topk_cnt"], z["topk_cnt"][0])
    if exact_scores:   # the oracle's NumPy arithmetic: equal to the last bit but for one ulp in ~1 of 2000 scores
        # (scipy evaluates exp() on an array, the restatement on a scalar: NumPy's two code paths differ there)
        assert np.allclose(res["topk_score"], z["topk_score"][0], rtol=1e-15, atol=0, equal_nan=True)
        assert np.allclose(res["topk2_score"], z["topk2_score"][0], rtol=1e-15, atol=0, equal_nan=True)
    else:   # device exp() vs NumPy's: 1e-5 is the contract (BASELINE.json), observed ~1e-19 absolute
        assert np.allclose(res["topk_score"], z["topk_score"][0], rtol=1e-12, atol=0, equal_nan=True)
        assert np.allclose(res["topk2_score"], z["topk2_score"][0], rtol=1e-12, atol=0, equal_nan=True)
    assert np.array_equal(res["mis_rank"], z["mis_rank"][0])
    assert np.array_equal(res["assign"], z["assign"])
    assert np.array_equal(res["n_cand"], z["per_span_candidates"])
    assert res["not_best_count"] == m["not_best_count"]
    assert res["cnt_unassigned"] == m["cnt_unassigned"]


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_equals_reference(path):
    g = Golden(path)
    prob, wins_before = _inputs(g)
    res = osk.solve_skip(prob.in_start, prob.in_end, prob.out_start, prob.out_end, prob.preds,
                         time_windows_before=wins_before)
    assert res["windows"] == [tuple(w) for w in g.meta["windows"]]
    _check_against_golden(g, res, exact_scores=True)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_host_mirror_tally_equals_reference(path):
    """time windows + WaterFill of traceweaver_b200.skipmode (NumPy) against the reference's."""
    g = Golden(path)
    prob, wins_before = _inputs(g)
    st = skipmode.SkipState()
    st.time_windows = list(wins_before)
    sorted_start = [np.sort(np.asarray(o, np.int64), kind="stable") for o in prob.out_start]
    wins, budgets, counts = skipmode.tally(prob.in_start, prob.in_end, sorted_start, st)
    assert [tuple(w) for w in g.meta["time_windows"]] == wins
    assert [g.meta["skip_budget"][ep] for ep in g.topo] == budgets
    assert np.array_equal(np.asarray([g.meta["skip_count"][ep] for ep in g.topo]), counts)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_device_code_stepped_on_cpu_equals_reference(path):
    """The kernel body of k_skip (tw_skip_core.cuh, compiled for the CPU by tests/emul) on the fixtures;
    time windows, skip counts and the BuildDistributions table come from the oracle here."""
    import emul_backend
    from oracle import tw_oracle
    g = Golden(path)
    prob, wins_before = _inputs(g)
    ref = osk.solve_skip(prob.in_start, prob.in_end, prob.out_start, prob.out_end, prob.preds,
                         time_windows_before=wins_before)
    res = emul_backend.skip_solve(prob.in_start, prob.in_end, prob.out_start, prob.out_end, prob.preds,
                                  ref["time_windows"], ref["skip_count"], ref["pair_params"], ref["skip_budget"])
    assert tw_oracle.windows_from_cuts(res["cut"]) == [tuple(w) for w in g.meta["windows"]]
    res = dict(res, topk2_idx=res["top2_idx"], topk2_score=res["top2_score"],
               not_best_count=int(res["counters"][0, 0]), cnt_unassigned=int(res["counters"][0, 1]),
               time_windows=ref["time_windows"], skip_budget=ref["skip_budget"], skip_count=ref["skip_count"],
               pair_params=ref["pair_params"], large_delay=ref["large_delay"])
    _check_against_golden(g, res, exact_scores=False)


def test_norm_pdf_restatement_matches_scipy():
    import scipy.stats
    rng = np.random.default_rng(3)
    for _ in range(200):
        x, loc, sc = rng.normal(0, 3000), rng.normal(500, 300), abs(rng.normal(800, 500)) + 1e-3
        assert osk.norm_pdf(x, loc, sc) == float(scipy.stats.norm.pdf(x, loc=loc, scale=sc))
        assert osk.norm_logpdf(x, loc, sc) == float(scipy.stats.norm.logpdf(x, loc=loc, scale=sc))


def test_exact_mwis_prefers_true_optimum_below_float_noise():
    # two in-spans, two ranks each; (rank 0, rank 0) and (rank 1, rank 1) are the independent pairs and
    # their totals differ by 3e-10 in favour of the second (the size of the gap in hotel_load150_cache20,
    # window 797-798): a tolerance of 1e-9 would call that a tie and return the first
    a = [(0.00029014, (10, 20)), (0.00029010, (10, 21))]
    b = [(0.00030727, (11, 21)), (0.00030731 + 3e-10, (11, 20))]
    assert osk.exact_mwis([a, b]) == [1, 1]
    b[1] = (0.00030731 - 3e-10, (11, 20))
    assert osk.exact_mwis([a, b]) == [0, 0]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_engine_equals_reference(path):
    from traceweaver_b200.engine import Engine
    g = Golden(path)
    prob, wins_before = _inputs(g)
    eng = Engine(0)
    st = skipmode.SkipState()
    st.time_windows = list(wins_before)
    res = skipmode.solve(eng, prob.in_start, prob.in_end, prob.out_start, prob.out_end, prob.preds,
                         labels=[g.meta["in_ep"]] + g.topo, state=st)
    eng.close()
    # PerfectCut flags -> the reference's window list
    from oracle import tw_oracle
    assert tw_oracle.windows_from_cuts(res["cut"]) == [tuple(w) for w in g.meta["windows"]]
    res = dict(res, topk2_idx=res["top2_idx"], topk2_score=res["top2_score"],
               not_best_count=int(res["counters"][0, 0]), cnt_unassigned=int(res["counters"][0, 1]))
    _check_against_golden(g, res, exact_scores=False)


def synthetic_two_skip_eps(seed=5, n=240):
    """A service with cache hits on TWO of three parallel callees (skip spans at two tuple positions,
    which no shipped dataset has; the third callee is complete, so no tuple is all-skip — on that the
    reference itself raises)."""
    rng = np.random.default_rng(seed)
    in_s = np.cumsum(rng.integers(2000, 9000, n)).astype(np.int64) + 1_600_000_000_000_000
    in_e = in_s + rng.integers(9000, 16000, n)
    outs_s, outs_e = [], []
    for e, missing in enumerate((30, 17, 0)):
        s = in_s + rng.integers(200, 3000, n)
        en = s + rng.integers(1000, 4000, n)
        keep = np.sort(rng.choice(n, n - missing, replace=False))
        s, en = s[keep], en[keep]
        od = np.argsort(s, kind="stable")
        outs_s.append(s[od])
        outs_e.append(en[od])
    return in_s, in_e, outs_s, outs_e, [[], [], []]


def test_device_code_on_cpu_equals_oracle_on_two_skip_eps():
    import emul_backend
    in_s, in_e, outs_s, outs_e, preds = synthetic_two_skip_eps()
    ref = osk.solve_skip(in_s, in_e, outs_s, outs_e, preds)
    res = emul_backend.skip_solve(in_s, in_e, outs_s, outs_e, preds, ref["time_windows"], ref["skip_count"],
                                  ref["pair_params"], ref["skip_budget"])
    for a, b in (("top2_idx", "topk2_idx"), ("topk_idx", "topk_idx"), ("mis_rank", "mis_rank"), ("assign", "assign")):
        assert np.array_equal(res[a], ref[b]), a
    assert ((res["assign"] == -2).sum(axis=1) > 0).tolist() == [True, True, False]


@pytest.mark.gpu
def test_engine_equals_oracle_on_synthetic_skips():
    from traceweaver_b200.engine import Engine
    in_s, in_e, outs_s, outs_e, preds = synthetic_two_skip_eps()
    ref = osk.solve_skip(in_s, in_e, outs_s, outs_e, preds)
    eng = Engine(0)
    res = skipmode.solve(eng, in_s, in_e, outs_s, outs_e, preds)
    eng.close()
    assert np.array_equal(res["skip_count"], np.asarray(ref["skip_count"]))
    assert np.array_equal(res["pair_params"], ref["pair_params"], equal_nan=True)
    assert np.array_equal(res["top2_idx"], ref["topk2_idx"])
    assert np.array_equal(res["topk_idx"], ref["topk_idx"])
    assert np.array_equal(res["mis_rank"], ref["mis_rank"])
    assert np.array_equal(res["assign"], ref["assign"])
    assert np.allclose(res["topk_score"], ref["topk_score"], rtol=1e-12, atol=0, equal_nan=True)
    assert (res["assign"] == -2).sum() > 0


@pytest.mark.gpu
def test_predictor_runs_a_cache_directory_like_the_reference():
    """The drop-in `TraceWeaverV3.FindAssignments` on the services of one cache-mode run in the executor's
    order (frontend with skip budgets first, then search without): ONE predictor instance carries the
    time windows / distribution samples from service to service like the reference's; the 6-tuples equal
    the reference's (assignments incl. ("Skip", "Skip"), top-K lists, counters, candidates per span)."""
    from test_gpu_pipeline import reference_call_args
    from traceweaver_b200.predictor import TraceWeaverV3
    files = sorted(glob.glob(os.path.join(CACHE_DIR, "hotel_load150_cache20__*.npz")))
    gs = {Golden(f).meta["process"]: Golden(f) for f in files}
    pred = TraceWeaverV3({}, {}, device=0)
    for process in ("frontend", "search"):
        g = gs[process]
        assert [tuple(w) for w in g.meta["time_windows_before"]] == [tuple(w) for w in pred.skip_state.time_windows]
        in_parts, out_parts, truth, G = reference_call_args(g)
        if process == "frontend":       # the fixture's truth holds -1 where the reference's dict says ('Skip', 'Skip')
            for e, ep in enumerate(g.topo):
                for i, j in enumerate(g.z["truth"][e]):
                    if j < 0:
                        truth[ep][(str(g.z["in_trace"][i]), str(g.z["in_sid"][i]))] = ("Skip", "Skip")
        a, topk, not_best, n, cands, unassigned = pred.FindAssignments(
            "MaxScoreBatchSubsetWithSkips", process, in_parts, out_parts, False, [], truth, G)
        z, m = g.z, g.meta
        assert (not_best, n, unassigned) == (m["not_best_count"], m["num_spans"], m["cnt_unassigned"])
        in_ids = [s.GetId() for s in list(in_parts.values())[0]]
        for e, ep in enumerate(g.topo):
            ids = [s.GetId() for s in out_parts[ep]]
            for i, iid in enumerate(in_ids):
                c = int(z["assign"][e, i])
                want = ids[c] if c >= 0 else (("NA", "NA") if c == -1 else ("Skip", "Skip"))
                assert a[ep][iid] == want, (process, ep, i)
                wl = [ids[c] if c >= 0 else ("Skip", "Skip") for c in z["topk_final"][i, :z["topk_final_cnt"][i], e]]
                assert topk[ep][iid] == wl, (process, ep, i)
        for i, iid in enumerate(in_ids):
            assert cands.get(iid, 0) == int(z["per_span_candidates"][i])
