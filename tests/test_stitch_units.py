"""The criterion behind the stitch kernel's units (csrc/tw_stitch.cu: k_stitch_units), checked on the
reference's own candidate sets: a perfect cut at in-span i is STRONG when, for every callee, the highest
candidate position of the in-spans before i lies below lower_bound(in_i.start) in that callee's list.
On every fixture the two sides of a strong cut must share no candidate span (so different warps may stitch
them in any order), and strong cuts must be frequent enough to be worth it."""
import numpy as np
import pytest

from golden_util import Golden, golden_files

FILES = golden_files()
IDS = [f.split("/")[-1][:-4] for f in FILES]


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_strong_cuts_separate_the_candidate_sets(path):
    g = Golden(path)
    z, prob = g.z, g.problem()
    n, E = prob.n_in, prob.E
    off, flat = z["pre_off"], z["pre_flat"]              # candidates_array of CreateWindows2 (V3:1041-1051)
    sets = [flat[off[i]:off[i + 1]] for i in range(n)]   # rows (ep in topological order, position)
    cut = np.zeros(n + 1, bool)
    for a, b in g.windows():                             # a window that ends before the last in-span ends at a cut
        cut[b + 1] = True
    first = np.stack([np.searchsorted(prob.out_start[e], prob.in_start, side="left") for e in range(E)])   # [E, n]
    hi = np.full((E, n), -1, np.int64)
    for i, s in enumerate(sets):
        for e in range(E):
            pos = s[s[:, 0] == e, 1]
            if len(pos):
                hi[e, i] = pos.max()
    run = np.maximum.accumulate(hi, axis=1)              # highest candidate position up to and including i
    strong = [i for i in range(1, n) if cut[i] and all(run[e, i - 1] < first[e, i] for e in range(E))]
    # every candidate of an in-span lies at or after lower_bound(in.start): the premise of the criterion
    for i, s in enumerate(sets):
        for e, p in s:
            assert p >= first[e, i]
    seen = [set() for _ in range(E)]                     # positions used before the current in-span
    strong_set = set(strong)
    before_cut = [set() for _ in range(E)]
    for i, s in enumerate(sets):
        if i in strong_set:
            before_cut = [set(x) for x in seen]
        for e, p in s:
            if before_cut[e]:
                assert int(p) not in before_cut[e], (i, e, p)     # nothing from before the last strong cut
            seen[int(e)].add(int(p))
    assert len(strong) >= (n // 200), "strong cuts too rare to split the service"
