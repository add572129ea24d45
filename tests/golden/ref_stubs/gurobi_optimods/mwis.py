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


def _bnb(n, adj, w, incumbent=-1.0):
    """Exact maximum-weight independent set by branch and bound, one connected component at a time.
    Bound: a greedy clique cover of the still-available vertices (the best an independent set can take
    from a clique is its heaviest vertex) — the candidates of one incoming span form a clique, so the
    cover finds them and the bound is tight enough for windows in which 30 incoming spans compete for
    interchangeable outgoing spans (the plain "sum of the remaining weights" bound needs hours there).
    `incumbent`: a known lower bound on the optimum (the MILP's value minus a hair) to prune with."""
    order_all = sorted(range(n), key=lambda v: -w[v])
    total_w, total_sol = 0.0, []
    seen = 0
    for seed in order_all:
        if (seen >> seed) & 1:
            continue
        comp, frontier = 1 << seed, 1 << seed
        while frontier:
            v = frontier.bit_length() - 1
            frontier &= ~(1 << v)
            nb = adj[v] & ~comp
            comp |= nb
            frontier |= nb
        seen |= comp
        members = [v for v in order_all if (comp >> v) & 1 and w[v] > 0]
        best = [-1.0, []]

        def bound(P):
            cliques = []          # [common-neighbourhood mask of the clique's members]
            ub = 0.0
            for v in members:
                if not (P >> v) & 1:
                    continue
                for k, common in enumerate(cliques):
                    if (common >> v) & 1:
                        cliques[k] = common & adj[v]
                        break
                else:
                    cliques.append(adj[v])
                    ub += w[v]          # members are visited by descending weight: the first is the heaviest
            return ub

        def rec(P, cur_w, chosen):
            if P == 0:
                if cur_w > best[0]:
                    best[0], best[1] = cur_w, list(chosen)
                return
            if cur_w + bound(P) <= best[0]:
                return
            for v in members:
                if (P >> v) & 1:
                    break
            chosen.append(v)
            rec(P & ~adj[v] & ~(1 << v), cur_w + w[v], chosen)
            chosen.pop()
            rec(P & ~(1 << v), cur_w, chosen)

        P0 = 0
        for v in members:
            P0 |= 1 << v
        rec(P0, 0.0, [])
        total_w += max(best[0], 0.0)
        total_sol += best[1]
    return total_w, sorted(total_sol)


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
