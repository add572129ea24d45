"""Per-phase cycle shares of the refit kernels (library built with -DTW_PROFILE_PHASES)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from traceweaver_b200 import _lib
_lib.SO_PATH = os.path.join(os.path.dirname(_lib.SO_PATH), "libtw_b200_prof.so")
import torch
from traceweaver_b200 import synth
from traceweaver_b200.batch import build_batch_from_blocks
from traceweaver_b200.engine import Engine
blocks = synth.hotel_stream(4096, 1000, seed=10); hb = build_batch_from_blocks(blocks)
eng = Engine(0); eng.bind(hb); eng.prepare()
p0 = eng.params_pass0(); sc = eng.score(p0, want_used=True); r0 = eng.stitch(p0, sc["cut"], undeleted=sc)
d, c = eng.delays(r0["assign"]); eng.gmm_refit(d, c)
lib = _lib.load()
buf = (C.c_ulonglong * 16)()
torch.cuda.synchronize(); lib.tw_debug_gmm_phases(buf, 1)
eng.gmm_refit(d, c); torch.cuda.synchronize()
lib.tw_debug_gmm_phases(buf, 1)
names = ["seeding", "lloyd", "init_mstep", "em", "score"]
tot = sum(buf[:5]) or 1
print(" ".join(f"{n}={100*buf[k]/tot:.1f}%" for k, n in enumerate(names)))
print(f"fits={buf[7]} lloyd_iters/fit={buf[5]/max(buf[7],1):.1f} em_iters/fit={buf[6]/max(buf[7],1):.1f}")
