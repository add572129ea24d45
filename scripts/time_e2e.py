"""End-to-end BatchSolver.solve time (host buffers in/out) against the number of service groups."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from traceweaver_b200 import synth
from traceweaver_b200.api import BatchSolver
from traceweaver_b200.batch import build_batch_from_blocks
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
blocks = synth.hotel_stream(S, 1000, seed=10); hb = build_batch_from_blocks(blocks)
for tok in (sys.argv[2:] or ["1", "2", "4", "8"]):          # "<groups>" or "<groups>:<first group fraction>"
    chunks = int(tok.split(":")[0])
    sv = BatchSolver(0, chunks=chunks)
    if ":" in tok:
        sv.FIRST_GROUP_FRACTION = float(tok.split(":")[1])
    for _ in range(3): sv.solve(hb)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): sv.solve(hb)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"groups={tok} e2e {dt*1e3:.2f} ms", flush=True)
    sv.close()
