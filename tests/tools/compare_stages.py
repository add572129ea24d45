"""Stage-by-stage comparison of the CUDA engine with the CPU oracle on a sample of a synthetic
workload (debugging aid; run on the GPU box).  TEST tooling: it imports oracle/.
Usage: python scripts/compare_stages.py {hotel,media,alibaba} [services] [n_in] [per_block]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle import tw_oracle
from traceweaver_b200 import shard
from traceweaver_b200.batch import build_batch_from_blocks
from traceweaver_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "media"
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2046
n_in = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
per = int(sys.argv[4]) if len(sys.argv) > 4 else 2

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import sample_blocks  # noqa: E402

sp = shard.stream_spec(wl, ns, n_in, 10)
blocks = shard.generate_slice(sp, 0, ns)
sample = sample_blocks(blocks, per)
hb = build_batch_from_blocks(sample)
P = hb.n_problems
print(f"{wl}: {P} services in the sample; blocks: {[b.name for b in sample]}")
ob = tw_oracle.OracleBatch(hb)
eng = Engine(0)
eng.bind(hb)
eng.prepare()


def cmp(name, a, b, per_prob_off=None, tol=None):
    a = a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = np.asarray(b)
    a = a.reshape(b.shape)
    if tol is None:
        bad = a != b
        if a.dtype.kind == "f":
            bad &= ~(np.isnan(a) & np.isnan(b))
    else:
        bad = ~(np.isclose(a, b, rtol=0, atol=tol) | (np.isnan(a) & np.isnan(b)))
    nbad = int(bad.sum())
    msg = f"  {name:24s} {'OK' if nbad == 0 else 'DIFF'}  ({nbad} of {bad.size})"
    if nbad and per_prob_off is not None:
        flat = np.flatnonzero(bad.reshape(bad.shape[0], -1).any(axis=1)) if bad.ndim > 1 else np.flatnonzero(bad)
        probs = np.searchsorted(per_prob_off, flat, side="right") - 1
        u, c = np.unique(probs, return_counts=True)
        msg += "  problems: " + ", ".join(f"{int(p)}({sample_name(int(p))}):{int(k)}" for p, k in list(zip(u, c))[:12])
        first = int(flat[0])
        msg += f"  first at {first} (problem {int(probs[0])}, local {first - int(per_prob_off[int(probs[0])])})"
    print(msg)
    return nbad


def sample_name(p):
    acc = 0
    for b in sample:
        acc += b.in_start.shape[0]
        if p < acc:
            return b.name
    return "?"


in_off = hb.prob_in_off
tup_off = hb.prob_tuple_off
# pass-0 parameters
p0 = eng.params_pass0()
g0 = ob.params_pass0()
cmp("params0", p0.table, g0)
# score pass 0
sc = eng.score(p0, want_used=True)
so = ob.score(gauss=g0)
cmp("cut", sc["cut"], so["cut"], in_off)
cmp("n_feasible", sc["n_feasible"], so["n_feasible"], in_off)
cmp("topk_cnt(p0)", sc["topk_cnt"], so["topk_cnt"], in_off)
cmp("topk_idx(p0)", sc["topk_idx"], so["topk_idx"], tup_off * 5)
cmp("topk_score(p0) 1e-9", sc["topk_score"], so["topk_score"], in_off, tol=1e-9)
print("  tiles redone:", eng.redo_tile_count(), "of", eng.tile_count())
# stitch pass 0: fast path and slow path against the oracle
st_o = ob.stitch(so["cut"], gauss=g0)
for label, und in (("fast", sc), ("slow", None)):
    r0 = eng.stitch(p0, sc["cut"], undeleted=und)
    torch.cuda.synchronize()
    print(f" stitch pass 0 ({label} path)")
    cmp("assign", r0["assign"], st_o["assign"], tup_off)
    cmp("mis_rank", r0["mis_rank"], st_o["mis_rank"], in_off)
    cmp("n_cand", r0["n_cand"], st_o["n_cand"], in_off)
    c = r0["counters"].cpu().numpy()
    cmp("counters[:, :3]", c[:, :3], st_o["counters"][:, :3])
    print("  engine status per problem:", np.unique(c[:, 3], return_counts=True), " max nodes engine/oracle",
          c[:, 2].max(), st_o["counters"][:, 2].max())
try:
    eng.status()
except Exception as ex:
    print("engine status:", ex)
# whole path
from traceweaver_b200.predictor import solve_bound
try:
    res = solve_bound(eng, seed_select=10)
    ref = tw_oracle.find_assignments(hb, 10, threads=16)
    print(" whole path")
    cmp("assign", res["assign"], ref["assign"], tup_off)
    cmp("mis_rank", res["mis_rank"], ref["mis_rank"], in_off)
    nb = cmp("topk_idx", res["topk_idx"], ref["topk_idx"], tup_off * 5)
    if nb:
        a = res["topk_idx"].cpu().numpy(); b = ref["topk_idx"]
        sa = res["topk_score"].cpu().numpy(); sb = ref["topk_score"]
        flat = np.flatnonzero(a != b)
        shown = 0
        seen = set()
        for f in flat:
            pr = int(np.searchsorted(tup_off * 5, f, side="right") - 1)
            E = int(hb.prob_ep_off[pr + 1] - hb.prob_ep_off[pr])
            i = (int(f) - int(tup_off[pr]) * 5) // (5 * E)
            if (pr, i) in seen:
                continue
            seen.add((pr, i))
            gi = int(in_off[pr]) + i
            o = 5 * (int(tup_off[pr]) + i * E)
            print(f"   problem {pr} ({sample_name(pr)}) in-span {i}: engine idx {a[o:o+5*E].reshape(5, E).tolist()} scores {sa[gi].tolist()}")
            print(f"   {' ' * 40} oracle idx {b[o:o+5*E].reshape(5, E).tolist()} scores {sb[gi].tolist()}")
            shown += 1
            if shown >= 4:
                break
    cmp("n_cand", res["n_cand"], ref["n_cand_total"], in_off)
    cmp("mix 1e-6", res["params_pass1"].table, ref["mix"], tol=1e-6)
except Exception as ex:
    print("whole path failed:", repr(ex)[:400])
