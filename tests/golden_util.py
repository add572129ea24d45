"""Load the golden fixtures minted by tests/golden/make_goldens.py."""
import glob
import json
import os

import numpy as np

from traceweaver_b200 import _abi
from traceweaver_b200.batch import Problem

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(pattern="*__*.npz", gpu=False):
    """Fixture files.  gpu=True leaves out the datasets listed in golden/gpu_unverified.txt: fixtures
    minted after the round's GPU budget was spent, checked on the CPU side only (oracle, emulation,
    loader, accuracy) until a GPU run has confirmed them."""
    files = sorted(glob.glob(os.path.join(GOLDEN_DIR, pattern)))
    if gpu:
        skip_file = os.path.join(GOLDEN_DIR, "gpu_unverified.txt")
        if os.path.exists(skip_file):
            skip = [ln.strip() for ln in open(skip_file) if ln.strip() and not ln.startswith("#")]
            files = [f for f in files if not any(os.path.basename(f).startswith(d + "__") for d in skip)]
    return files


class Golden:
    def __init__(self, path):
        self.path = path
        self.z = np.load(path)
        self.meta = json.loads(str(self.z["meta"]))
        m = self.meta
        self.name = f"{m['dataset']}/{m['process']}"
        given = m["out_eps_given"]
        self.topo = m["out_eps_topo"]
        self.pos_given = [given.index(ep) for ep in self.topo]     # topo position -> given index
        self.E = len(self.topo)

    def problem(self) -> Problem:
        z, m = self.z, self.meta
        topo = self.topo
        pos = {ep: i for i, ep in enumerate(topo)}
        outs_s, outs_e = [], []
        for g in self.pos_given:
            s = z[f"out{g}_start"].astype(np.int64)
            outs_s.append(s)
            outs_e.append(s + z[f"out{g}_dur"].astype(np.int64))
        preds = [[pos[b] for b in m["graph_in_edges"][ep]] for ep in topo]
        in_s = z["in_start"].astype(np.int64)
        return Problem(in_start=in_s, in_end=in_s + z["in_dur"].astype(np.int64), out_start=outs_s,
                       out_end=outs_e, preds=preds, name=self.name)

    def term_keys(self, prob: Problem):
        """(ep1, ep2) services_times key of every term, in term order."""
        in_ep = self.meta["in_ep"]
        keys = []
        for e, src in prob.terms():
            if src >= 0:
                keys.append((self.topo[src], self.topo[e]))
            elif src == _abi.TW_TERM_ROOT:
                keys.append((in_ep, self.topo[e]))
            else:
                keys.append((self.topo[e], in_ep))
        return keys

    def gauss_table(self, prob: Problem):
        """[n_batches, n_terms, 3] from the recorded ComputeEpPairDistParams3 snapshots."""
        keys = self.term_keys(prob)
        snaps = self.meta["params_pass0"]
        tab = np.zeros((len(snaps), len(keys), _abi.TW_GAUSS_REC))
        for bi, snap in enumerate(snaps):
            assert snap["start"] == bi * _abi.TW_PARAM_BATCH
            for t, k in enumerate(keys):
                mu, sd = snap["params"]["|".join(k)]
                if sd < 1e-12:
                    sd = 0.001
                tab[bi, t] = (mu, sd, np.log(sd))
        return tab

    def mix_table(self, prob: Problem):
        keys = self.term_keys(prob)
        tab = np.zeros((len(keys), _abi.TW_MIX_REC))
        for t, k in enumerate(keys):
            g = self.meta["params_pass1"]["|".join(k)]
            tab[t] = mix_record(g["weights"], g["means"], g["precisions_cholesky"])
        return tab

    def windows(self):
        return [tuple(w) for w in self.meta["windows"]]


def mix_record(weights, means, precisions_chol):
    """TW_MIX_REC layout: k, pc[5], mu*pc[5], log(pc)[5], log(w)[5]."""
    w = np.asarray(weights, np.float64)
    mu = np.asarray(means, np.float64)
    pc = np.asarray(precisions_chol, np.float64)
    k = len(w)
    rec = np.zeros(_abi.TW_MIX_REC)
    rec[0] = k
    rec[1:1 + k] = pc
    rec[6:6 + k] = mu * pc
    rec[11:11 + k] = np.log(pc)
    rec[16:16 + k] = np.log(w)
    return rec
