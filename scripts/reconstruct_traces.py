"""Reconstruct a directory of Jaeger JSON traces end to end: loader -> batch engine -> accuracy.

    python scripts/reconstruct_traces.py <trace dir> [--layout hotel|media|node|alibaba] [--device 0]

Prints, per solved service, the assignment accuracy against the traces' own parent links
(the reference's AccuracyForService, helpers/utils.py:34-60) and the time of each stage — the same
numbers executor.py prints for `--predictor_indices 10`."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("directory")
ap.add_argument("--layout", default="hotel", choices=["hotel", "media", "node", "alibaba"])
ap.add_argument("--device", type=int, default=0)
ap.add_argument("--seed", type=int, default=10)
args = ap.parse_args()

from traceweaver_b200.api import BatchSolver
from traceweaver_b200.loader import load_jaeger_dir, to_host_batch, accuracy

t0 = time.perf_counter()
from traceweaver_b200.engine import Engine
_eng = Engine(args.device)
services = load_jaeger_dir(args.directory, layout=args.layout, engine=_eng)    # truth + FindOrder on the device
_eng.close()
# services with n_out == n_in at every callee go through the two-pass batch path; a raw trace directory holds
# no others (skip budgets come from executor.py's cache transform, see traceweaver_b200.skipmode)
ok = [s for s in services if all(len(o) == s.problem.n_in for o in s.problem.out_start)]
t1 = time.perf_counter()
print(f"loaded {len(services)} services ({len(ok)} without skip budgets) in {t1 - t0:.2f} s")
hb = to_host_batch(ok)
solver = BatchSolver(device=args.device, seed_select=args.seed)
solver.solve(hb)                                   # warm-up: allocations, random streams
t2 = time.perf_counter()
out = solver.solve(hb)
t3 = time.perf_counter()
n_spans = sum(s.problem.n_in * (1 + s.problem.E) for s in ok)
print(f"solved {n_spans} spans in {1e3 * (t3 - t2):.2f} ms (host buffers in and out)")
for p, s in enumerate(ok):
    a = out["assign"][int(hb.prob_tuple_off[p]):int(hb.prob_tuple_off[p + 1])]
    print(f"  {s.name:28s} n_in={s.problem.n_in:5d} E={s.problem.E}  accuracy {100 * accuracy(s, a):7.3f} %")
