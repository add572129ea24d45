#!/usr/bin/env python
"""Mint golden vectors by running the UNMODIFIED reference in this container.

The reference (read-only at /root/reference) has no tests or known-answer vectors,
so parity is pinned by executing its own `executor.py` (predictor index 10 =
"MaxScoreBatchSubsetWithSkips" -> TraceWeaverV3, executor.py:899) on its shipped Jaeger JSON
datasets and recording, per solved service:

  * the problem exactly as handed to `TraceWeaverV3.FindAssignments` (executor.py:1172-1175):
    in/out span partitions (ids, start_mus, duration_mus), the invocation graph (node order,
    edge order, per-node in_edges order) and the ground truth,
  * every intermediate the engine must reproduce: topological ep order, perfect-cut windows
    (traceweaver_v3.py:1020-1078), the Gaussian parameters of every 100-span batch
    (:580-646), the fitted GMMs (:706-818), per in-span top-K lists with and without deletion
    (:1182,:1185), the MWIS choice per window (:1193) and
  * the 6-tuple it returns (:1229) plus the accuracies executor.py prints.

How the reference is made to run here: stub modules for packages that are absent
(tests/golden/ref_stubs: matplotlib, deepdiff, pygmmis, gurobi_optimods -> exact MWIS), a
writable copy of the dataset directory (the shipped time_order_filenames.pickle holds the
author's absolute paths), and ONE in-memory source substitution in traceweaver_v3.py:790
(`true_durations != []` on an ndarray raises under numpy>=2; debug print only).  A second
version-skew shim: the reference calls `GaussianMixture.score` with an INTEGER sample
(traceweaver_v1.py:126); scikit-learn 1.9 (this container) allocates its log-probability buffer
with the input's dtype and silently truncates it to int64, whereas the pinned scikit-learn 1.5.1
(requirements.txt:22) computes in float64.  The harness casts the sample to float64 before
calling sklearn so the goldens carry the pinned version's (documented) semantics.  Nothing is
written under /root/reference and no reference source is copied into this repository.

NumPy's global RNG is re-seeded (np.random.seed(GLOBAL_SEED)) immediately before each
FindAssignments call: the reference's BIC model selection (traceweaver_v3.py:774) draws from
the unseeded global RNG, so without this the goldens would not be reproducible.  The product's
"sklearn" refit backend documents and applies the same convention.

Usage:  python tests/golden/make_goldens.py hotel_load100 [node_load100 media_load100 ...]
Outputs tests/golden/<dataset>__<service>.npz (+ <dataset>.json with the printed accuracies).
This script cannot run on the GPU box (/root/reference is absent there); its outputs are
committed.
"""
import io
import json
import os
import runpy
import shutil
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
PORT = os.path.join(REF, "src/trace_reconstructor/ports/python")
GLOBAL_SEED = 10

DATASETS = {
    # name: (relative data dir, --fix, test_name)   (exps/exp1/run_experiment.sh:5-62)
    **{f"hotel_load{l}": (f"data/hotel_reservation/hotel_load{l}", 2) for l in (25, 50, 75, 100, 125, 150)},
    **{f"media_load{l}": (f"data/media_microservices/media_load{l}", 1) for l in (25, 50, 75, 100, 125, 150)},
    **{f"node_load{l}": (f"data/nodejs_microservices/node_load{l}", 0) for l in (25, 50, 75, 100, 125, 150)},
    # the Alibaba JSON layout (--fix 5) on synthetic traces of that layout (tests/golden/make_alibaba_traces.py;
    # the real trace is a Git LFS pointer): absolute path, not under /root/reference
    "alibaba_synth": (os.path.join(HERE, "alibaba_synth"), 5),
}


def _install_reference_modules():
    sys.path.insert(0, PORT)
    sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
    import algorithms  # namespace package rooted at PORT/algorithms
    path = os.path.join(PORT, "algorithms/traceweaver_v3.py")
    src = open(path).read()
    needle = "iteration == 1 and true_durations != []:"
    assert src.count(needle) == 1
    src = src.replace(needle, "iteration == 1 and False:")
    mod = types.ModuleType("algorithms.traceweaver_v3")
    mod.__file__ = path
    mod.__package__ = "algorithms"
    sys.modules["algorithms.traceweaver_v3"] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    algorithms.traceweaver_v3 = mod
    _shim_sklearn_int_score()
    return mod


def _shim_sklearn_int_score():
    """sklearn>=1.6 truncates log-probabilities to int64 for integer input (see module doc)."""
    from sklearn import mixture
    orig = mixture.GaussianMixture.score
    if getattr(orig, "_tw_shim", False):
        return

    def score(self, X, y=None):
        return orig(self, np.asarray(X, dtype=np.float64), y)

    score._tw_shim = True
    mixture.GaussianMixture.score = score


class Recorder:
    """Wraps TraceWeaverV3 methods; one record per FindAssignments call."""

    def __init__(self, v3mod, slice_in_spans=None):
        self.v3 = v3mod
        self.records = []
        self.cur = None
        self.slice_in_spans = slice_in_spans
        self.dump_as = None
        cls = v3mod.TraceWeaverV3
        self.orig = {n: getattr(cls, n) for n in (
            "FindAssignments", "FindTopKAssignments", "GetAssignmentsMIS", "CreateWindows2",
            "ComputeEpPairDistParams3", "ComputeEpPairDistParams5", "TallySkipSpans", "BuildDistributions")}
        rec = self

        def FindAssignments(self_, method, process, in_parts, out_parts, parallel, hops, truth, graph, *a, **k):
            return rec.find_assignments(self_, method, process, in_parts, out_parts, parallel, hops, truth, graph, *a, **k)

        def FindTopKAssignments(self_, in_eps, in_span, out_eps, parts, K, graph, id_x, preprocess_phase=False, count_candidates_phase=True):
            res = rec.orig["FindTopKAssignments"](self_, in_eps, in_span, out_eps, parts, K, graph, id_x, preprocess_phase, count_candidates_phase)
            rec.on_topk(in_span, out_eps, K, preprocess_phase, count_candidates_phase, res)
            return res

        def GetAssignmentsMIS(self_, top_assignments):
            res = rec.orig["GetAssignmentsMIS"](self_, top_assignments)
            rec.on_mis(top_assignments, res)
            return res

        def CreateWindows2(self_, *a, **k):
            res = rec.orig["CreateWindows2"](self_, *a, **k)
            rec.cur["windows"] = [tuple(w) for w in res]
            return res

        def ComputeEpPairDistParams3(self_, in_parts, out_parts, out_eps, s, e, graph):
            res = rec.orig["ComputeEpPairDistParams3"](self_, in_parts, out_parts, out_eps, s, e, graph)
            rec.on_params3(self_, s)
            return res

        def ComputeEpPairDistParams5(self_, in_parts, out_parts, graph, all_assignments, truth):
            res = rec.orig["ComputeEpPairDistParams5"](self_, in_parts, out_parts, graph, all_assignments, truth)
            rec.on_params5(self_)
            return res

        def TallySkipSpans(self_, in_parts, out_parts, in_eps, out_eps, batch_size_mis):
            res = rec.orig["TallySkipSpans"](self_, in_parts, out_parts, in_eps, out_eps, batch_size_mis)
            rec.on_tally(self_, out_eps)
            return res

        def BuildDistributions(self_, process, in_parts, out_parts, in_eps, out_eps):
            res = rec.orig["BuildDistributions"](self_, process, in_parts, out_parts, in_eps, out_eps)
            rec.cur["build_dist"] = rec._snap(self_.services_times)
            rec.cur["large_delay"] = int(self_.large_delay)
            return res

        for n, f in list(locals().items()):
            if n in self.orig:
                setattr(cls, n, f)

    # -- helpers -----------------------------------------------------------------------
    def find_assignments(self, inst, method, process, in_parts, out_parts, parallel, hops, truth, graph, *a, **k):
        assert method == "MaxScoreBatchSubsetWithSkips"
        in_ep, in_spans = list(in_parts.items())[0]
        only, skip = os.environ.get("TW_GOLDEN_ONLY"), os.environ.get("TW_GOLDEN_SKIP")
        if (only and process not in only.split(",")) or (skip and process in skip.split(",")):
            # service filtered out of this minting run (the slow services are minted by their own
            # process): hand the executor an all-"NA" answer of the right shape, record nothing
            ids = [s.GetId() for s in in_spans]
            return ({ep: {i: ("NA", "NA") for i in ids} for ep in out_parts},
                    {ep: {i: [] for i in ids} for ep in out_parts}, 0, len(ids), {i: 0 for i in ids}, len(ids))
        cur = self.cur = {
            "process": process, "in_ep": in_ep,
            "in_ids": [s.GetId() for s in in_spans],
            "in_start": [s.start_mus for s in in_spans],
            "in_dur": [s.duration_mus for s in in_spans],
            "out_eps_given": list(out_parts.keys()),
            "out_ids": {ep: [s.GetId() for s in p] for ep, p in out_parts.items()},
            "out_start": {ep: [s.start_mus for s in p] for ep, p in out_parts.items()},
            "out_dur": {ep: [s.duration_mus for s in p] for ep, p in out_parts.items()},
            "graph_nodes": list(graph.nodes()),
            "graph_edges": list(graph.edges()),
            "graph_in_edges": {n: [u for u, _ in graph.in_edges(n)] for n in graph.nodes()},
            "truth": {ep: dict(d) for ep, d in truth.items()},
            "topk": [], "topk2": [], "pre": [], "mis": [], "params3": [], "params5": [],
            "iteration_marks": [],
        }
        cur["id2idx"] = {ep: {sid: i for i, sid in enumerate(ids)} for ep, ids in cur["out_ids"].items()}
        # instance state the reference carries from one service to the next (V3:35-48 never resets it)
        cur["time_windows_before"] = [tuple(w) for w in inst.time_windows]
        cur["dist_values_before"] = {"|".join(k): list(map(float, v)) for k, v in inst.distribution_values.items()}
        cur["dynamism_before"] = bool(inst.dynamism)
        cur["skip_code"] = {}
        np.random.seed(GLOBAL_SEED)
        t0 = time.time()
        res = self.orig["FindAssignments"](inst, method, process, in_parts, out_parts, parallel, hops, truth, graph, *a, **k)
        cur["seconds"] = time.time() - t0
        cur["result"] = res
        cur["out_eps_topo"] = list(inst.GetOutEpsInOrder(out_parts, graph))
        self.records.append(cur)
        self.cur = None
        if self.dump_as:          # write the fixture as soon as the service is done
            print("minted", _dump(self.dump_as, cur, getattr(self, "dump_dir", HERE)), "%.0fs" % cur["seconds"], file=sys.__stdout__, flush=True)
        return res

    def _tuple_idx(self, out_eps, spans):
        idx = []
        for ep, s in zip(out_eps, spans[1:]):
            if s.trace_id == "None":          # a skip span (V3:983-995): -2 - (its index among the ep's skip spans)
                idx.append(-2 - self.cur["skip_code"][(ep, s.sid)])
                continue
            idx.append(self.cur["id2idx"][ep][s.GetId()])
        return idx

    def on_topk(self, in_span, out_eps, K, pre, count, res):
        cur = self.cur
        if pre:
            # K=-1 unscored enumeration: entries are stacks [(ep, span), ...]
            ids = set()
            for stack in res:
                for ep, s in stack[1:]:
                    ids.add((ep, cur["id2idx"][ep][s.GetId()]))
            cur["pre"].append((len(res), sorted(ids)))
            return
        entry = [(float(score), self._tuple_idx(out_eps, spans)) for score, spans in res]
        (cur["topk"] if count else cur["topk2"]).append(entry)

    def on_tally(self, inst, out_eps):
        """After TallySkipSpans (V3:853-989): the time windows in FetchSkipFromWindow's order, the
        water-filled skip counts, and an index for every skip span (windows in sorted order, then
        position in the window's list) so that tuples can name them."""
        cur = self.cur
        wins = sorted(inst.time_windows, key=lambda x: x[0])
        cur["time_windows"] = [tuple(w) for w in wins]
        cur["skip_budget"] = {ep: int(inst.overall_skip_budget[ep]) for ep in out_eps}
        cur["skip_count"] = {}
        for ep in out_eps:
            counts, g = [], 0
            for w in wins:
                lst = inst.available_skips_per_window[ep][tuple(w[:2])]
                counts.append(len(lst))
                for pos, (sp, used) in enumerate(lst):
                    assert used == 0
                    cur["skip_code"].setdefault((ep, sp.sid), g + pos)
                g += len(lst)
            cur["skip_count"][ep] = counts

    def on_mis(self, top_assignments, res):
        chosen = []
        for cand, a in zip(top_assignments, res):
            r = -1
            for j, (score, spans) in enumerate(cand):
                if spans is a:
                    r = j
            if len(a) > 0:
                assert r >= 0
            chosen.append(r)
        self.cur["mis"].append(chosen)

    @staticmethod
    def _snap(times):
        out = {}
        for key, v in times.items():
            if isinstance(v, tuple):
                out[key] = ("gauss", float(v[0]), float(v[1]))
            elif hasattr(v, "weights_"):
                out[key] = ("gmm", v.weights_.copy(), v.means_.reshape(-1).copy(),
                            v.covariances_.reshape(-1).copy(), v.precisions_cholesky_.reshape(-1).copy())
            else:
                out[key] = ("other", repr(type(v)))
        return out

    def on_params3(self, inst, start):
        self.cur["params3"].append((start, self._snap(inst.services_times)))

    def on_params5(self, inst):
        # called at the end of EVERY iteration (v3:1221-1222); the snapshot after iteration 0 is
        # what iteration 1 scores with, the one after iteration 1 is never used
        self.cur["params5"].append(self._snap(inst.services_times))


def _dump(dataset, rec, outdir):
    process = rec["process"]
    out_eps = rec["out_eps_topo"]
    E = len(out_eps)
    n = len(rec["in_ids"])
    K = 5
    all_assign, all_topk, not_best, n_spans, per_span_cand, cnt_un = rec["result"]
    id2idx = rec["id2idx"]
    passes = len(rec["topk"]) // n
    assert len(rec["topk"]) == passes * n and len(rec["topk2"]) == passes * n and len(rec["pre"]) == n

    def pack_topk(entries):
        sc = np.full((passes, n, K), np.nan)
        ix = np.full((passes, n, K, E), -1, np.int32)
        cnt = np.zeros((passes, n), np.int32)
        for t, entry in enumerate(entries):
            p, i = divmod(t, n)
            cnt[p, i] = len(entry)
            for r, (s, idx) in enumerate(entry):
                sc[p, i, r] = s
                ix[p, i, r] = idx
        return sc, ix, cnt

    tk_s, tk_i, tk_c = pack_topk(rec["topk"])
    t2_s, t2_i, t2_c = pack_topk(rec["topk2"])

    assign = np.full((E, n), -1, np.int32)
    for e, ep in enumerate(out_eps):
        for i, iid in enumerate(rec["in_ids"]):
            v = all_assign[ep][iid]
            if v == ("NA", "NA"):
                assign[e, i] = -1
            elif v == ("Skip", "Skip"):
                assign[e, i] = -2
            else:
                assign[e, i] = id2idx[ep][v]
    truth = np.full((E, n), -1, np.int32)
    for e, ep in enumerate(out_eps):
        for i, iid in enumerate(rec["in_ids"]):
            v = rec["truth"][ep].get(iid)
            truth[e, i] = id2idx[ep][v] if v in id2idx[ep] else -1
    topk_final = np.full((n, K, E), -1, np.int32)
    topk_final_cnt = np.zeros(n, np.int32)
    for e, ep in enumerate(out_eps):
        for i, iid in enumerate(rec["in_ids"]):
            lst = all_topk[ep][iid]
            topk_final_cnt[i] = len(lst)
            for r, v in enumerate(lst):
                topk_final[i, r, e] = -2 if v == ("Skip", "Skip") else id2idx[ep][v]

    # MWIS choice: one chosen rank (or -1) per in-span and pass, in window order
    mis = np.full((passes, n), -9, np.int32)
    wins = rec["windows"]
    for p in range(passes):
        for w, (ws, we) in enumerate(wins):
            chosen = rec["mis"][p * len(wins) + w]
            assert len(chosen) == we - ws + 1
            mis[p, ws:we + 1] = chosen
    assert (mis != -9).all()

    pre_cnt = np.array([c for c, _ in rec["pre"]], np.int32)
    pre_ids = [ids for _, ids in rec["pre"]]
    pre_off = np.zeros(n + 1, np.int64)
    for i, ids in enumerate(pre_ids):
        pre_off[i + 1] = pre_off[i] + len(ids)
    ep_pos = {ep: e for e, ep in enumerate(out_eps)}
    pre_flat = np.array([(ep_pos[ep], j) for ids in pre_ids for ep, j in ids], np.int32).reshape(-1, 2)

    # parameter tables: pass 0 -> one snapshot per 100-span batch; pass 1 -> GMMs
    in_ep = rec["in_ep"]
    keys = []
    preds = rec["graph_in_edges"]
    for ep in rec["out_eps_given"]:
        if len(preds[ep]) == 0:
            keys.append((in_ep, ep))
        for b in preds[ep]:
            keys.append((b, ep))  # non-primary keys are filtered below if absent
        keys.append((ep, in_ep))
    p3 = []
    for start, snap in rec["params3"]:
        row = {}
        for k in keys:
            if k in snap and snap[k][0] == "gauss":
                row["|".join(k)] = [snap[k][1], snap[k][2]]
        p3.append({"start": int(start), "params": row})
    p5 = {}
    if rec["params5"]:
        snap5 = rec["params5"][0]
        for k in keys:
            if k in snap5 and snap5[k][0] == "gmm":
                _, w, m, c, pc = snap5[k]
                p5["|".join(k)] = {"weights": w.tolist(), "means": m.tolist(), "covariances": c.tolist(),
                                   "precisions_cholesky": pc.tolist()}

    meta = {
        "dataset": dataset, "process": process, "in_ep": in_ep,
        "out_eps_given": rec["out_eps_given"], "out_eps_topo": out_eps,
        "graph_nodes": rec["graph_nodes"], "graph_edges": rec["graph_edges"],
        "graph_in_edges": rec["graph_in_edges"],
        "windows": rec["windows"], "passes": passes,
        "not_best_count": int(not_best), "num_spans": int(n_spans), "cnt_unassigned": int(cnt_un),
        "params_pass0": p3, "params_pass1": p5,
        "reference_seconds": rec["seconds"], "global_seed": GLOBAL_SEED,
        # skip / cache mode (row f-4): state carried into the call, TallySkipSpans' and BuildDistributions' results
        "time_windows_before": rec.get("time_windows_before", []), "time_windows": rec.get("time_windows", []),
        "dynamism_before": rec.get("dynamism_before", False),
        "dist_values_before_keys": sorted(rec.get("dist_values_before", {}).keys()),
        "skip_budget": rec.get("skip_budget", {}), "skip_count": rec.get("skip_count", {}),
        "large_delay": rec.get("large_delay"),
        "build_dist": {"|".join(k): [v[1], v[2]] for k, v in rec.get("build_dist", {}).items() if v[0] == "gauss"},
        "versions": _versions(),
    }
    arrays = {
        "in_start": np.array(rec["in_start"], np.int64), "in_dur": np.array(rec["in_dur"], np.int64),
        "in_trace": np.array([a for a, _ in rec["in_ids"]]), "in_sid": np.array([b for _, b in rec["in_ids"]]),
        "assign": assign, "truth": truth, "topk_final": topk_final, "topk_final_cnt": topk_final_cnt,
        "topk_score": tk_s, "topk_idx": tk_i, "topk_cnt": tk_c,
        "topk2_score": t2_s, "topk2_idx": t2_i, "topk2_cnt": t2_c,
        "mis_rank": mis, "pre_cnt": pre_cnt, "pre_off": pre_off, "pre_flat": pre_flat,
        "per_span_candidates": np.array([per_span_cand.get(iid, 0) for iid in rec["in_ids"]], np.int64),
        "meta": np.array(json.dumps(meta)),
    }
    for e, ep in enumerate(rec["out_eps_given"]):
        arrays[f"out{e}_start"] = np.array(rec["out_start"][ep], np.int64)
        arrays[f"out{e}_dur"] = np.array(rec["out_dur"][ep], np.int64)
        arrays[f"out{e}_trace"] = np.array([a for a, _ in rec["out_ids"][ep]])
        arrays[f"out{e}_sid"] = np.array([b for _, b in rec["out_ids"][ep]])
    safe = process.replace("/", "_").replace(" ", "_")
    path = os.path.join(outdir, f"{dataset}__{safe}.npz")
    np.savez_compressed(path, **arrays)
    return path


def _versions():
    import networkx, scipy, sklearn
    return {"python": sys.version.split()[0], "numpy": np.__version__, "scipy": scipy.__version__,
            "sklearn": sklearn.__version__, "networkx": networkx.__version__,
            "mwis": "exact (HiGHS mip_rel_gap=0 cross-checked with B&B); Gurobi unavailable"}


def run_dataset(name, outdir, v3mod, recorder):
    # "hotel_load150@0.2": the dataset with --cache_rate 0.2 (exps/exp2/run_experiment.sh); fixtures go
    # to tests/golden_cache/ under the name hotel_load150_cache20
    cache_rate = "0"
    if "@" in name:
        name, cache_rate = name.split("@")
        outdir = os.path.join(os.path.dirname(HERE), "golden_cache")
        os.makedirs(outdir, exist_ok=True)
    label = name if cache_rate == "0" else f"{name}_cache{int(round(float(cache_rate) * 100))}"
    rel, fix = DATASETS[name]
    scratch = tempfile.mkdtemp(prefix="tw_golden_")
    data = os.path.join(scratch, name)
    shutil.copytree(rel if os.path.isabs(rel) else os.path.join(REF, rel), data)
    cache = os.path.join(data, "time_order_filenames.pickle")
    if os.path.exists(cache):
        os.remove(cache)
    results = os.path.join(scratch, "results") + "/"
    os.makedirs(results)
    argv = ["executor.py", "--absolute_path", data, "--compressed", "0", "--cache_rate", cache_rate,
            "--fix", str(fix), "--test_name", name, "--load_level", "0", "--compress_factor", "1",
            "--repeat_factor", "1", "--execute_parallel", "0", "--results_directory", results,
            "--clear_cache", "0", "--predictor_indices", "10"]
    old_argv, old_cwd, old_stdout = sys.argv, os.getcwd(), sys.stdout
    sys.argv = argv
    os.chdir(scratch)
    buf = io.StringIO()
    recorder.records = []
    recorder.dump_as = label
    recorder.dump_dir = outdir
    t0 = time.time()
    try:
        sys.stdout = buf
        runpy.run_path(os.path.join(PORT, "executor.py"), run_name="__main__")
    finally:
        sys.stdout = old_stdout
        sys.argv = argv and old_argv
        os.chdir(old_cwd)
    wall = time.time() - t0
    log = buf.getvalue()
    acc_lines = [l for l in log.splitlines() if "ccuracy" in l and "iteration" not in l]
    paths = [_dump(label, rec, outdir) for rec in recorder.records]
    if os.environ.get("TW_GOLDEN_ONLY") or os.environ.get("TW_GOLDEN_SKIP"):
        return paths, {"printed_accuracy": [], "wall_seconds": wall}      # partial run: no summary file
    summary = {"dataset": label, "fix": fix, "cache_rate": float(cache_rate), "wall_seconds": wall,
               "find_assignments_seconds": {r["process"]: r["seconds"] for r in recorder.records},
               "printed_accuracy": acc_lines, "versions": _versions()}
    with open(os.path.join(outdir, f"{label}.json"), "w") as f:
        json.dump(summary, f, indent=1)
    shutil.rmtree(scratch, ignore_errors=True)
    return paths, summary


def main():
    names = sys.argv[1:] or ["hotel_load100"]
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    v3mod = _install_reference_modules()
    recorder = Recorder(v3mod)
    for name in names:
        paths, summary = run_dataset(name, HERE, v3mod, recorder)
        print(name, "->", [os.path.basename(p) for p in paths])
        for l in summary["printed_accuracy"]:
            print("   ", l)
        print("    wall %.1fs" % summary["wall_seconds"])


if __name__ == "__main__":
    main()
