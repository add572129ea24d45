"""Per-phase cycle shares and work counters of k_score3 (library built with -DTW_PROFILE_PHASES:
python -m traceweaver_b200.csrc.build --prof).  Usage: score_phase_profile.py [n_services] [workload]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traceweaver_b200 import _lib
_lib.SO_PATH = os.path.join(os.path.dirname(_lib.SO_PATH), "libtw_b200_prof.so")
import torch
from traceweaver_b200 import shard
from traceweaver_b200.batch import build_batch_from_blocks
from traceweaver_b200.engine import Engine
from traceweaver_b200.predictor import solve_bound
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = sys.argv[2] if len(sys.argv) > 2 else "hotel"
sp = shard.stream_spec(wl, S, 1000, 10)
hb = build_batch_from_blocks(shard.generate_slice(sp, 0, S))
eng = Engine(0); eng.bind(hb)
res = solve_bound(eng)
lib = _lib.load()
buf = (C.c_ulonglong * 24)()
names = ["setup+staging", "ranges", "admission/rest", "term tables", "2a feasibility", "2b scores", "2c top-K", "results"]
n_in = int(hb.prob_in_off[-1])
for label, prm in (("gauss pass (windows + top-K)", None), ("mixture pass (top-K only)", res["params_pass1"])):
    torch.cuda.synchronize(); lib.tw_debug_score_phases(buf, 1)
    if prm is None:
        p0 = eng.params_pass0(); eng.score(p0, want_used=True)
    else:
        eng.score(prm, out=dict(cut=res["cut"]), keep_windows=True)
    torch.cuda.synchronize(); lib.tw_debug_score_phases(buf, 1)
    tot = sum(buf[:8]) or 1
    print(label)
    print("  " + " ".join(f"{n}={100*buf[k]/tot:.1f}%" for k, n in enumerate(names)))
    rounds = max(buf[12], 1)
    print(f"  per in-span: slots {buf[10]/n_in:.2f} (valid {buf[14]/n_in:.2f}) combos {buf[11]/n_in:.2f} feasible {buf[13]/n_in:.2f}; "
          f"rounds per warp-tile {rounds/ (n_in/32):.2f}; warp cycles per in-span {tot/n_in:.0f}")
    print(f"  tiles flagged for the sequential kernel: staging overflow {buf[15]}, range/combination limits {buf[16]} warps, "
          f"ties/NaN {buf[17]} warps; tiles {eng.tile_count()}")
