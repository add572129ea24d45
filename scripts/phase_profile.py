"""Per-phase cycle shares of k_score2 (library built with -DTW_PROFILE_PHASES)."""
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
d, c = eng.delays(r0["assign"]); p1 = eng.gmm_refit(d, c)
lib = _lib.load()
names = ["load_view", "stage+params", "ranges+helper", "scan+admit(+serial)", "slots", "combos", "select", "finalize", "cut+maps"]
for label, prm in (("gauss", p0), ("gmm", p1)):
    buf = (C.c_ulonglong * 16)()
    torch.cuda.synchronize(); lib.tw_debug_score2_phases(buf, 1)
    eng.score(prm, want_used=True); torch.cuda.synchronize()
    lib.tw_debug_score2_phases(buf, 1)
    tot = sum(buf[:8]) or 1
    print(label, " ".join(f"{n}={100*buf[k]/tot:.1f}%" for k, n in enumerate(names[:8])))
