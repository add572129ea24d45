"""Exact maximum-weight-independent-set stand-in for gurobi_optimods.mwis (reference call site
traceweaver_v3.py:1411).  Two independent exact solvers are run and must agree:
  * HiGHS through scipy.optimize.milp with mip_rel_gap=0 (one x_u+x_v<=1 row per edge),
  * a plain branch-and-bound in Python.
The B&B answer is returned (vertex indices, ascending), so the golden vectors are defined as
"the exact MWIS optimum", not as whatever a MIP gap tolerance lets through.
Test infrastructure only."""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import milp, LinearConstraint, Bounds

STATS = {"solves": 0, "max_vertices": 0, "max_edges": 0, "disagree": 0}
RECORD = []          # optional: (n, edges, weights, solution) tuples for oracle cross-checks
RECORD_ENABLED = False


def _bnb(n, adj, w):
    order = sorted(range(n), key=lambda v: -w[v])
    best = [-1.0, []]
    suffix = [0.0] * (n + 1)
    for k in range(n - 1, -1, -1):
        suffix[k] = suffix[k + 1] + max(w[order[k]], 0.0)

    def rec(k, cur_w, chosen, banned):
        if cur_w > best[0]:
            best[0] = cur_w
            best[1] = list(chosen)
        if k == n or cur_w + suffix[k] <= best[0]:
            return
        v = order[k]
        if not (banned >> v) & 1 and w[v] > 0:
            chosen.append(v)
            rec(k + 1, cur_w + w[v], chosen, banned | adj[v])
            chosen.pop()
        rec(k + 1, cur_w, chosen, banned)

    rec(0, 0.0, [], 0)
    return best[0], sorted(best[1])


def maximum_weighted_independent_set(adjacency_matrix, weights, verbose=False):
    A = sp.coo_array(adjacency_matrix)
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    edges = [(int(u), int(v)) for u, v, d in zip(A.row, A.col, A.data) if d != 0 and u != v]
    adj = [0] * n
    for u, v in edges:
        adj[u] |= 1 << v
        adj[v] |= 1 << u
    bw, bsol = _bnb(n, adj, [float(x) for x in w])

    STATS["solves"] += 1
    STATS["max_vertices"] = max(STATS["max_vertices"], n)
    STATS["max_edges"] = max(STATS["max_edges"], len(edges))

    if edges:
        rows = np.repeat(np.arange(len(edges)), 2)
        cols = np.array(edges).reshape(-1)
        C = sp.csr_array((np.ones(len(rows)), (rows, cols)), shape=(len(edges), n))
        res = milp(c=-w, constraints=LinearConstraint(C, -np.inf, 1.0), integrality=np.ones(n),
                   bounds=Bounds(0, 1), options={"mip_rel_gap": 0.0})
        assert res.status == 0, res.message
        hsol = sorted(int(i) for i in np.flatnonzero(res.x > 0.5))
        hw = float(w[hsol].sum())
        if abs(hw - bw) > 1e-7:
            STATS["disagree"] += 1
            raise AssertionError(f"exact MWIS solvers disagree: highs {hw} vs bnb {bw}")
    if RECORD_ENABLED:
        RECORD.append((n, edges, w.copy(), list(bsol)))
    return np.array(bsol, dtype=np.int64)
