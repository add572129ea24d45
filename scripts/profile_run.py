"""One pass of the whole path on a small synthetic stream — the command ncu wraps.
Usage: python scripts/profile_run.py [n_services]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from traceweaver_b200 import synth
from traceweaver_b200.batch import build_batch_from_blocks
from traceweaver_b200.engine import Engine
from traceweaver_b200.predictor import solve_bound
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
if len(sys.argv) > 2:      # bench.py's streams: "media" / "alibaba" / "hotel"
    from traceweaver_b200 import shard
    blocks = shard.generate_slice(shard.stream_spec(sys.argv[2], S, 1000, 10), 0, S)
else:
    blocks = synth.hotel_stream(S, 1000, seed=10)
hb = build_batch_from_blocks(blocks)
eng = Engine(0)
eng.bind(hb)
solve_bound(eng)   # warm-up pass (skipped by ncu -s)
solve_bound(eng)
torch.cuda.synchronize()
print("done", synth.span_count(blocks))
