"""Evidence for tests/test_oracle_gmm.py:ILL_CONDITIONED — scikit-learn's own BIC on a nodejs delay
sample (7 distinct values) depends on the ORDER of the samples: same multiset, same starting point,
different summation order inside `resp.T @ (X * X)`.

    python tests/gmm_conditioning.py [tests/golden/node_load50__service2.npz] [term]
"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # test tooling: may use oracle/
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import numpy as np
from sklearn import mixture
from golden_util import Golden
from oracle import tw_oracle
from traceweaver_b200.batch import build_batch

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests/golden/node_load50__service2.npz")
term = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = Golden(path)
prob = g.problem(); hb = build_batch([prob]); ob = tw_oracle.OracleBatch(hb)
n, E = prob.n_in, prob.E
assign0 = np.full((E, n), -1, np.int32)
for i in range(n):
    if g.z["mis_rank"][0][i] >= 0:
        assign0[:, i] = g.z["topk_idx"][0][i, g.z["mis_rank"][0][i]]
d, c = ob.delays(assign0.reshape(-1)); off = hb.term_sample_off
x = d[off[term]:off[term] + c[term]].astype(float)
vals, cnts = np.unique(x, return_counts=True)
print("distinct delays:", dict(zip(vals.astype(int).tolist(), cnts.tolist())))
rng = np.random.default_rng(0)
for k in (3, 4, 5):
    init = mixture.GaussianMixture(n_components=k, covariance_type="diag", random_state=3, max_iter=1).fit(x.reshape(-1, 1))
    bics = []
    for trial in range(6):
        xp = x if trial == 0 else rng.permutation(x)
        gm = mixture.GaussianMixture(n_components=k, covariance_type="diag", weights_init=init.weights_,
                                     means_init=init.means_, precisions_init=init.precisions_).fit(xp.reshape(-1, 1))
        bics.append(gm.bic(xp.reshape(-1, 1)))
    print(f"k={k}: BIC over 6 orderings of the same samples: min {min(bics):.3f} max {max(bics):.3f} spread {max(bics) - min(bics):.3f}")
