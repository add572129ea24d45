"""GPU-vs-oracle diff of the skip path on one fixture (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from golden_util import Golden
from oracle import tw_oracle_skip as osk
from oracle import tw_oracle
from traceweaver_b200 import skipmode
from traceweaver_b200.engine import Engine
g = Golden(sys.argv[1])
prob = g.problem()
eng = Engine(0)
st = skipmode.SkipState()
res = skipmode.solve(eng, prob.in_start, prob.in_end, prob.out_start, prob.out_end, prob.preds,
                     labels=[g.meta["in_ep"]] + g.topo, state=st)
z = g.z
gw = [tuple(w) for w in g.meta["windows"]]
cut_ref = np.zeros(len(prob.in_start), np.uint8)
for (a, b) in gw:
    pass
ow = tw_oracle.windows_from_cuts(res["cut"])
print("windows equal", ow == gw, len(ow), len(gw))
for a, b in zip(ow, gw):
    if a != b:
        print("first window diff", a, b); break
print("cut sum", int(res["cut"].sum()))
for k, zk in (("top2_idx", "topk2_idx"), ("topk_idx", "topk_idx")):
    d = np.flatnonzero((res[k] != z[zk][0]).any(axis=(1, 2)))
    print(k, "rows differing", len(d), d[:10])
print("assign diff", int((res["assign"] != z["assign"]).sum()), "mis diff", int((res["mis_rank"] != z["mis_rank"][0]).sum()))
print("pair ok", all(res["pair_params"][([g.meta["in_ep"]] + g.topo).index(k.split("|")[0]), ([g.meta["in_ep"]] + g.topo).index(k.split("|")[1])][0] == v[0]
                     for k, v in g.meta["build_dist"].items() if k.split("|")[0] in [g.meta["in_ep"]] + g.topo and k.split("|")[1] in [g.meta["in_ep"]] + g.topo))
print("skip_count ok", np.array_equal(np.asarray([g.meta["skip_count"][ep] for ep in g.topo]), res["skip_count"]))
print("score maxdiff", np.nanmax(np.abs(res["topk_score"] - z["topk_score"][0])), np.nanmax(np.abs(res["top2_score"] - z["topk2_score"][0])))
