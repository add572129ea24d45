"""Latency of the drop-in plugin call (TraceWeaverV3.FindAssignments signature, Span objects in,
dicts out) on the hotel fixtures, next to the reference's own time for the same call
(recorded when the goldens were minted).  Usage: python scripts/dropin_latency.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import Golden, golden_files
from test_gpu_pipeline import reference_call_args
from traceweaver_b200.predictor import TraceWeaverV3

pred = TraceWeaverV3({}, {}, device=0)
tot_ref = tot_us = 0.0
for f in golden_files(gpu=True):
    g = Golden(f)
    args = reference_call_args(g)
    for _ in range(2):   # second call: steady state (allocations, caches warm)
        in_parts, out_parts, truth, G = args
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pred.FindAssignments("MaxScoreBatchSubsetWithSkips", g.meta["process"], in_parts, out_parts, False, [], truth, G)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n_spans = g.problem().n_in * (1 + g.E)
    tot_ref += g.meta["reference_seconds"]; tot_us += dt
    print(f"{g.name:36s} {n_spans:6d} spans  reference {g.meta['reference_seconds']:7.2f} s   plugin call {dt*1e3:8.2f} ms   x{g.meta['reference_seconds']/dt:8.0f}")
print(f"total: reference {tot_ref:.1f} s, plugin {tot_us*1e3:.1f} ms, x{tot_ref/tot_us:.0f}")
