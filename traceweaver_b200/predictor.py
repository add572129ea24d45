"""Drop-in predictor: the reference's plugin interface for the accelerated path.

Mirrors `TraceWeaverV3` of the reference (src/trace_reconstructor/ports/python/algorithms/
traceweaver_v3.py:29, constructed as `Cls(all_spans, all_processes)` and registered in
executor.py:888-900) for method "MaxScoreBatchSubsetWithSkips": same constructor, same
`FindAssignments` signature (traceweaver_v3.py:1087), same 6-tuple result (:1229), so the
reference's executor.py and helpers/utils.py accuracy code run unmodified on top of it (see
INTEGRATION.md).  All compute is in the CUDA engine behind the C ABI; this file only marshals
`Span` objects to SoA arrays and index results back to span ids.
"""
import numpy as np
import torch

from . import _abi, refit, skipmode
from .batch import Problem, build_batch
from .engine import Engine

METHOD = "MaxScoreBatchSubsetWithSkips"
NA = ("NA", "NA")
SKIP = ("Skip", "Skip")


def solve_batch(eng: Engine, hb, seed_select=10, truth_assign=None, term_order=None, device_arrays=None,
                pinned=False, keep=("assign", "mis_rank", "counters", "n_cand", "topk")):
    """The whole path for a bound-able batch: both passes of traceweaver_v3.py:1159-1227.

    params0 -> windows -> stitch (pass 0) -> delays -> refit -> score (final top-K) -> stitch
    (pass 1).  Returns device tensors; one host sync at the end (engine status)."""
    eng.bind(hb, device_arrays=device_arrays, pinned=pinned)
    return solve_bound(eng, seed_select, truth_assign, term_order)


def solve_bound(eng: Engine, seed_select=10, truth_assign=None, term_order=None, check=True, after_score=None):
    """Both passes over the batch currently bound to `eng` (inputs resident in HBM).  check=False
    leaves the one host sync (engine status) to the caller, who must call eng.status().
    `after_score(top)` is called as soon as the final top-K lists are queued (before the last
    stitch), so a caller can start copying them out while the stitch runs."""
    eng.prepare()                                 # prev-index scan, sorted end times
    p0 = eng.params_pass0()                       # ComputeEpPairDistParams3, every 100-span batch
    # CreateWindows2 (perfect-cut flags) + FindTopKAssignments on the undeleted lists, pass-0 params
    sc = eng.score(p0, want_used=True)
    r0 = eng.stitch(p0, sc["cut"], undeleted=sc)  # iteration 0
    delays, counts = eng.delays(r0["assign"])     # ComputeEpPairDistParams5: durations
    base = None
    if truth_assign is not None:
        # the reference fits (and discards) GMMs on the TRUE assignments first; they only advance
        # NumPy's global random stream
        dt, ct = eng.delays(truth_assign)
        base = eng.gmm_stream_draws(dt, ct)
    p1 = eng.gmm_refit(delays, counts, seed_select=seed_select, prob_base_skip=base, term_order=term_order)
    # top_k_2 of the last iteration -> all_topk_assignments; candidate maps are parameter independent
    top = eng.score(p1, out=dict(used_lo=sc["used_lo"], used_bits=sc["used_bits"], used_wide=sc["used_wide"],
                                 cut=sc["cut"]), keep_windows=True)
    if after_score is not None:
        after_score(top)
    r1 = eng.stitch(p1, sc["cut"], undeleted=top)  # iteration 1
    n_cand = r0["n_cand"] + r1["n_cand"]          # per_span_candidates accumulates over iterations
    if check:
        eng.status()
    return dict(assign=r1["assign"], mis_rank=r1["mis_rank"], counters=r1["counters"], n_cand=n_cand,
                topk_score=top["topk_score"], topk_idx=top["topk_idx"], topk_cnt=top["topk_cnt"],
                cut=sc["cut"], params_pass1=p1, assign_pass0=r0["assign"])


class TraceWeaverV3:
    """`predictors` entry replacing the reference's ("MaxScoreBatchSubsetWithSkips", TraceWeaverV3)."""

    def __init__(self, all_spans, all_processes, device=0, seed_select=10, carry_state=True):
        self.all_spans = all_spans
        self.all_processes = all_processes
        self.seed_select = seed_select
        self.engine = Engine(device)          # raises without a CUDA device: there is no CPU path
        self.last = None
        # what the reference's instance keeps from one service to the next and its skip regime reads
        # (time_windows, distribution_values: traceweaver_v3.py:40,45 are never reset)
        self.skip_state = skipmode.SkipState()
        self.carry_state = carry_state
        self._pending_dist = []

    # -- marshalling -------------------------------------------------------------------------------
    @staticmethod
    def _arrays(spans):
        raw = [sp.start_mus for sp in spans]
        if any(isinstance(x, float) and x != int(x) for x in raw):
            # executor.py --compress_factor > 1 (transforms.repeat_change_spans) divides the start times: float
            # microseconds, which the engine's int64 timestamps cannot hold — loud, not truncated
            raise NotImplementedError("fractional start_mus (time-compressed spans): outside the engine's int64 "
                                      "timestamp model; use the reference for --compress_factor > 1")
        s = np.fromiter((int(x) for x in raw), np.int64, len(spans))
        d = np.fromiter((sp.duration_mus for sp in spans), np.int64, len(spans))
        return s, s + d

    def _problem(self, process, in_spans, out_span_partitions, invocation_graph):
        import networkx as nx
        out_eps = list(nx.topological_sort(invocation_graph))            # traceweaver_v1.py:37-39
        if set(out_eps) != set(out_span_partitions.keys()):
            raise ValueError("invocation_graph nodes must be the outgoing endpoints")
        pos = {ep: i for i, ep in enumerate(out_eps)}
        in_s, in_e = self._arrays(in_spans)
        outs = [self._arrays(out_span_partitions[ep]) for ep in out_eps]
        preds = [[pos[b] for b, _ in invocation_graph.in_edges(ep)] for ep in out_eps]
        prob = Problem(in_start=in_s, in_end=in_e, out_start=[o[0] for o in outs], out_end=[o[1] for o in outs],
                       preds=preds, name=process)
        return prob, out_eps

    # -- the reference's entry point ---------------------------------------------------------------
    def FindAssignments(self, method, process, in_span_partitions, out_span_partitions, parallel,
                        instrumented_hops, true_assignments, invocation_graph, true_skips=False,
                        true_dist=False):
        assert len(in_span_partitions) == 1                              # traceweaver_v3.py:1088
        if method != METHOD or true_skips or true_dist:
            raise NotImplementedError(
                f"traceweaver_b200 accelerates method {METHOD!r} only (got {method!r}); the ablation "
                "variants stay with the reference implementation")
        in_ep, in_spans = list(in_span_partitions.items())[0]
        # TallySkipSpans re-sorts every partition by start (stable), traceweaver_v3.py:968-971
        in_spans = sorted(in_spans, key=lambda x: float(x.start_mus))
        if any(len(p) != len(in_spans) for p in out_span_partitions.values()):
            # skip budgets (cache hits / dynamism): ONE iteration with skip spans, traceweaver_v3.py:1138-1158
            return self._find_assignments_skip(process, in_ep, in_spans, out_span_partitions, true_assignments,
                                               invocation_graph)
        out_parts = {ep: sorted(p, key=lambda x: float(x.start_mus)) for ep, p in out_span_partitions.items()}
        prob, out_eps = self._problem(process, in_spans, out_parts, invocation_graph)
        hb = build_batch([prob])
        n, E = prob.n_in, prob.E

        in_ids = [s.GetId() for s in in_spans]
        out_ids = [[s.GetId() for s in out_parts[ep]] for ep in out_eps]
        # ground truth only advances the refit's random stream, exactly as in the reference
        truth = np.full((E, n), -1, np.int32)
        for e, ep in enumerate(out_eps):
            lut = {sid: j for j, sid in enumerate(out_ids[e])}
            ta = true_assignments.get(ep, {}) if true_assignments else {}
            truth[e] = [lut.get(ta.get(i), -1) for i in in_ids]
        given_pos = [out_eps.index(ep) for ep in out_span_partitions.keys()]
        order = np.asarray(refit.reference_term_order(prob, given_pos), np.int32)

        if self.carry_state:
            # the reference runs TallySkipSpans and BuildDistributions for EVERY service (v3:1136, :1149);
            # in this regime they only leave state behind for a later service with skip budgets: the time
            # windows are appended now (a few tuples), the distribution samples are derived lazily — the
            # arrays are parked and run through tw_build_dist_samples, in call order, when a service with
            # skip budgets actually arrives (_find_assignments_skip)
            self.skip_state.time_windows.extend(skipmode.new_time_windows(prob.in_start, prob.in_end))
            self._pending_dist.append((prob.in_start, prob.in_end, prob.out_start, prob.out_end, [in_ep] + out_eps))
        dev = self.engine.device
        res = solve_batch(self.engine, hb, seed_select=self.seed_select,
                          truth_assign=torch.from_numpy(truth.reshape(-1)).to(dev),
                          term_order=torch.from_numpy(order).to(dev))
        self.last = res
        assign = res["assign"].cpu().numpy().reshape(E, n)
        topk_idx = res["topk_idx"].cpu().numpy().reshape(n, _abi.TW_K, E)
        topk_cnt = res["topk_cnt"].cpu().numpy()
        n_cand = res["n_cand"].cpu().numpy()
        counters = res["counters"].cpu().numpy()

        all_assignments, all_topk = {}, {}
        for e, ep in enumerate(out_eps):
            ids = out_ids[e]
            col = assign[e]
            all_assignments[ep] = {in_ids[i]: (ids[col[i]] if col[i] >= 0 else NA) for i in range(n)}
            all_topk[ep] = {in_ids[i]: [ids[topk_idx[i, r, e]] for r in range(topk_cnt[i])] for i in range(n)}
        per_span_candidates = {}
        for ep in out_span_partitions.keys():                            # traceweaver_v3.py:1096-1098
            for key in (true_assignments.get(ep, {}) if true_assignments else {}):
                per_span_candidates[key] = 0
        for i in range(n):
            if n_cand[i] or in_ids[i] in per_span_candidates:
                per_span_candidates[in_ids[i]] = int(n_cand[i])
        return (all_assignments, all_topk, int(counters[0, 0]), n, per_span_candidates, int(counters[0, 1]))

    # -- skip / cache mode --------------------------------------------------
    def _find_assignments_skip(self, process, in_ep, in_spans, out_span_partitions, true_assignments,
                               invocation_graph):
        import networkx as nx
        out_eps = list(nx.topological_sort(invocation_graph))            # traceweaver_v1.py:37-39
        if set(out_eps) != set(out_span_partitions.keys()):
            raise ValueError("invocation_graph nodes must be the outgoing endpoints")
        pos = {ep: i for i, ep in enumerate(out_eps)}
        in_s, in_e = self._arrays(in_spans)
        outs = [self._arrays(out_span_partitions[ep]) for ep in out_eps]     # the caller's list order
        preds = [[pos[b] for b, _ in invocation_graph.in_edges(ep)] for ep in out_eps]
        state = self.skip_state if self.carry_state else skipmode.SkipState()
        for a in self._pending_dist:                  # BuildDistributions of the earlier services, in call order
            skipmode.build_distributions(self.engine, *a[:4], a[4], state)
        self._pending_dist = []
        res = skipmode.solve(self.engine, in_s, in_e, [o[0] for o in outs], [o[1] for o in outs], preds,
                             labels=[in_ep] + out_eps, state=state, want_topk=False)
        self.last = res
        n, E = len(in_spans), len(out_eps)
        in_ids = [s.GetId() for s in in_spans]
        out_ids = [[s.GetId() for s in out_span_partitions[ep]] for ep in out_eps]

        def name(e, c):
            return out_ids[e][c] if c >= 0 else (NA if c == -1 else SKIP)    # traceweaver_v1.py:446-453
        assign, top2, cnt = res["assign"], res["top2_idx"], res["top2_cnt"]
        all_assignments = {ep: {in_ids[i]: name(e, int(assign[e, i])) for i in range(n)} for e, ep in enumerate(out_eps)}
        all_topk = {ep: {in_ids[i]: [name(e, int(top2[i, r, e])) for r in range(cnt[i])] for i in range(n)}
                    for e, ep in enumerate(out_eps)}
        per_span_candidates = {}
        for ep in out_span_partitions.keys():                            # traceweaver_v3.py:1096-1098
            for key in (true_assignments.get(ep, {}) if true_assignments else {}):
                per_span_candidates[key] = 0
        n_cand = res["n_cand"]
        for i in range(n):
            if n_cand[i] or in_ids[i] in per_span_candidates:
                per_span_candidates[in_ids[i]] = int(n_cand[i])
        ctr = res["counters"]
        return (all_assignments, all_topk, int(ctr[0, 0]), n, per_span_candidates, int(ctr[0, 1]))
