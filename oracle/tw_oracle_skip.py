"""CPU restatement of the reference's skip / cache mode.

TEST INFRASTRUCTURE ONLY — the checker for `tw_skip_pass`, never the thing measured or shipped.
Pinned to golden vectors minted from the reference itself with `--cache_rate` > 0
(tests/golden/make_goldens.py hotel_load150@0.2 ..., fixtures under tests/golden_cache/).

Literal, list-based restatement (pure-Python loops: the cases are one service of 1000 in-spans) of
what `TraceWeaverV3.FindAssignments` (traceweaver_v3.py = V3, traceweaver_v1.py = V1) does when some
outgoing endpoint has a skip budget != 0 (V3:1138-1158: `dynamism`, ONE iteration):

  TallySkipSpans / WaterFill / TackleMismatch   V3:853-989
  BuildDistributions                             V3:108-172
  FetchSkipFromWindow                            V3:820-842
  FindTopKAssignments / DfsTraverseX skip branch V3:219-234, :292-351
  ScoreAssignmentAsPerInvocationGraph            V1:259-361 (FindValidAncestor :264-292, normalized)
  GetEpPairCost                                  V1:117-139 (scipy.stats.norm.pdf / logpdf restated)
  BuildMISInstance + exact MWIS                  V3:1252-1281, :1395-1419
  AddAssignment / AddTopKAssignments             V1:433-488

Spans are (start, end) integer pairs; an out span is named by its index in its ep's list, a skip
span by the code -2 - g, g = its index among the ep's skip spans (time windows in
FetchSkipFromWindow's sorted order, then position in the window's list).  Labels: 0 = the incoming
endpoint, 1 + e = outgoing endpoint e (topological position).
"""
import bisect
import heapq
import math

import numpy as np

K = 5
MAX_WINDOW = 30
WEIGHT_OFFSET = 10000.0
SQRT_2PI = math.sqrt(2 * math.pi)          # scipy.stats._continuous_distns._norm_pdf_C
LOG_SQRT_2PI = math.log(2 * math.pi) / 2.0  # _norm_pdf_logC


class ReferenceUndefined(Exception):
    """The reference itself raises on this input (score tie between a skip span and a real span,
    all-skip tuple, ancestor chain of skips, missing distribution key)."""


def norm_pdf(x, loc, scale):
    """scipy.stats.norm.pdf(x, loc, scale): exp(-y**2 / 2) / sqrt(2 pi) / scale, y = (x - loc) / scale."""
    y = (x - loc) / scale
    return float(np.exp(-y ** 2 / 2.0) / SQRT_2PI / scale)


def norm_logpdf(x, loc, scale):
    """scipy.stats.norm.logpdf: -y**2 / 2 - log(sqrt(2 pi)) - log(scale)."""
    y = (x - loc) / scale
    return float(-y ** 2 / 2.0 - LOG_SQRT_2PI - np.log(scale))


# ---------------------------------------------------------------------------------------------------
# TallySkipSpans, V3:853-989
# ---------------------------------------------------------------------------------------------------
def new_time_windows(in_start, in_end):
    """The (start, end, 30) windows TallySkipSpans appends to self.time_windows, V3:973-985."""
    n = len(in_start)
    wins = []
    ws = int(in_start[0])
    final_end = int(max(in_end))
    for i in range(n):
        if i != 0 and i != n - 1 and i % MAX_WINDOW == 0:
            we = int(in_end[i])
            wins.append((ws, we, MAX_WINDOW))
            ws = we
        elif i == n - 1:
            wins.append((ws, final_end, MAX_WINDOW))
    return wins


def water_fill(existing, expected, budget):
    """WaterFill, V3:863-917, for one ep.  existing/expected: per window in SORTED-window order.
    Tie order among equal counts is numpy's argsort, as in the reference (V3:884)."""
    num = len(existing)
    alloc = np.zeros(num)
    if budget <= 0:
        return alloc
    existing = np.asarray(existing, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    order = np.argsort(existing)[::-1]
    srt = existing[order]
    lam = 0
    total_remaining = 0
    for i in range(num):
        lam = (budget + np.sum(srt[:i + 1])) // (i + 1)
        total_remaining = (budget + np.sum(srt[:i + 1])) % (i + 1)
        if lam <= srt[i]:
            break
    remaining = 0
    for i in range(num):
        want = max(lam - srt[i], 0)
        got = min(want, expected[i] - srt[i])          # (sic) expected is indexed by sorted position
        remaining += want - got
        alloc[order[i]] = got
    total_remaining += remaining
    while total_remaining > 0:
        no_change = True
        for i in reversed(range(num)):
            if total_remaining > 0 and alloc[order[i]] < (expected[i] - srt[i]):
                alloc[order[i]] += 1
                no_change = False
                total_remaining -= 1
        if no_change:
            break
    return alloc


def tally_skip_spans(in_start, in_end, out_start, time_windows_before):
    """Returns (windows sorted by start, budgets per ep, skip counts [E][n_windows])."""
    n = len(in_start)
    wins_all = list(time_windows_before) + new_time_windows(in_start, in_end)
    wins = sorted(wins_all, key=lambda w: w[0])
    budgets = [n - len(o) for o in out_start]
    counts = []
    for e, o in enumerate(out_start):
        existing = []
        for (ws, we, _) in wins:
            # spans with ws < start <= we (V3:935-937)
            existing.append(bisect.bisect_right(o, we) - bisect.bisect_right(o, ws))
        alloc = water_fill(existing, [w[2] for w in wins], budgets[e])
        counts.append([max(int(a), 0) for a in alloc])     # range(int(count)) of a negative count is empty
    return wins, budgets, counts


# ---------------------------------------------------------------------------------------------------
# BuildDistributions, V3:108-172
# ---------------------------------------------------------------------------------------------------
def build_distribution_samples(in_start, in_end, out_start, out_end):
    """The (key, value) samples one call appends to self.distribution_values, in order.
    key = (label of the parent's ep, label of the span's ep)."""
    spans = []          # (start, end, kind, label)
    for s, e in zip(in_start, in_end):
        spans.append((int(s), int(e), 0, 0))                       # server, incoming endpoint
    for ep, (os_, oe_) in enumerate(zip(out_start, out_end)):
        for s, e in zip(os_, oe_):
            spans.append((int(s), int(e), 1, 1 + ep))              # client, out ep (topological position)
    spans.sort(key=lambda x: x[0])                                 # stable
    large_delay = max(int(e) - int(s) for s, e in zip(in_start, in_end))
    samples = []
    for i, (s, e, kind, lab) in enumerate(spans):
        if kind == 1:
            parent = None
            for j in range(i - 1, -1, -1):
                ps, pe, pk, pl = spans[j]
                if e - ps > large_delay:
                    break
                if pk == 0:
                    parent = (pl, s - ps)
                    break
                if pk == 1 and pe < s and pl < lab:
                    parent = (pl, s - pe)
                    break
            if parent is not None:
                samples.append(((parent[0], lab), parent[1]))
        else:
            parent = None
            for j in range(i - 1, -1, -1):
                ps, pe, pk, pl = spans[j]
                if e - ps > large_delay:
                    break
                if pk == 1 and pe < e:
                    parent = (pl, e - pe)
                    break
            if parent is not None:
                samples.append(((parent[0], lab), parent[1]))
            samples.append(((lab, lab), e - s))
    return samples, large_delay


def pair_params(samples, E, values_before=None):
    """services_times after BuildDistributions as a dense [(E+1), (E+1), 2] table (NaN = key absent):
    np.mean / np.std over the accumulated lists (V3:171-172)."""
    lists = {}
    for k, v in (values_before or {}).items():
        lists[k] = list(v)
    for k, v in samples:
        lists.setdefault(k, []).append(v)
    tab = np.full((E + 1, E + 1, 2), np.nan)
    for (a, b), v in lists.items():
        tab[a, b, 0] = np.mean(v)
        tab[a, b, 1] = np.std(v)
    return tab


# ---------------------------------------------------------------------------------------------------
# candidate windows (CreateWindows2, V3:1020-1078) — enumeration WITH cutoffs, without skip spans
# ---------------------------------------------------------------------------------------------------
def _bisect_left(a, x):
    """bisect.bisect_left — also on a list that is NOT sorted (see solve_skip): the plain halving loop."""
    lo, hi = 0, len(a)
    while lo < hi:
        mid = (lo + hi) // 2
        if a[mid] < x:
            lo = mid + 1
        else:
            hi = mid
    return lo


def _bisect_right(a, x):
    lo, hi = 0, len(a)
    while lo < hi:
        mid = (lo + hi) // 2
        if x < a[mid]:
            hi = mid
        else:
            lo = mid + 1
    return lo


def feasible_sets(in_start, in_end, out_start, out_end, preds):
    """candidates_array of CreateWindows2 (V3:1041-1051): FindCutoffs (V3:182-217, literal: halving
    searches on the lists AS GIVEN, Python's negative-index wrap) + DfsTraverse3 (V3:236-288)."""
    E = len(out_start)
    succ = [[s for s in range(E) if e in preds[s]] for e in range(E)]
    res = []
    for s_in, e_in in zip(in_start, in_end):
        cut = [[len(o) - 1, 0] for o in out_start]
        for node in reversed(range(E)):
            exit_t = e_in
            for nb in succ[node]:
                exit_t = min(exit_t, out_start[nb][cut[nb][1]])       # index -1 wraps to the last span
            cut[node][0] = _bisect_left(out_start[node], s_in)
            cut[node][1] = _bisect_right(out_start[node], exit_t) - 1
        used = set()

        def rec(level, chosen):
            if level == E:
                for e, x in enumerate(chosen):
                    used.add((e, x))
                return
            os_, oe_ = out_start[level], out_end[level]
            for x in range(max(cut[level][0], 0), min(cut[level][1], len(os_) - 1) + 1):
                if s_in > os_[x] or oe_[x] > e_in:
                    continue
                if any(out_end[b][chosen[b]] > os_[x] for b in preds[level]):
                    continue
                rec(level + 1, chosen + [x])
        rec(0, [])
        res.append(used)
    return res


def windows_from_sets(in_end, sets):
    n = len(sets)
    prev_index = 0
    windows = []
    count = 1
    wstart = 0
    for i in range(n):
        if i != 0:
            if i == n - 1:
                count = 0
                windows.append((wstart, i))
            else:
                if i == 1:
                    prev_index = 0
                elif in_end[i - 1] >= in_end[prev_index]:
                    prev_index = i - 1
                cut = sets[prev_index].isdisjoint(sets[i]) and in_end[prev_index] <= in_end[i]
                if cut:
                    count = 0
                    windows.append((wstart, i - 1))
                    wstart = i
                elif count == MAX_WINDOW:
                    count = 0
                    windows.append((wstart, i))
                    wstart = i + 1
        else:
            wstart = i
        count += 1
    return windows


# ---------------------------------------------------------------------------------------------------
# the hot loop with skips
# ---------------------------------------------------------------------------------------------------
class _Entry:
    """(score, stack) heap entry with the reference's comparison: score first; on equal scores the
    first position holding two different spans decides by Span.__lt__ = start_mus (spans.py:51) —
    which raises TypeError when one of them is a skip span (start_mus == "None")."""
    __slots__ = ("score", "tup", "starts")

    def __init__(self, score, tup, starts):
        self.score, self.tup, self.starts = score, tup, starts

    def __lt__(self, other):
        if self.score != other.score:
            return self.score < other.score
        for a, b, sa, sb in zip(self.tup, other.tup, self.starts, other.starts):
            if a != b:
                if a < 0 or b < 0:
                    raise ReferenceUndefined("score tie between a skip span and a real span")
                return sa < sb
        return False


def solve_skip(in_start, in_end, out_start, out_end, preds, time_windows_before=(), values_before=None,
               mwis=None):
    """One call of FindAssignments in the skip regime.  preds[e] = predecessor positions in in_edges
    order.  Returns a dict of index arrays shaped like the golden fixtures.

    The out lists are taken AS THE CALLER HANDS THEM OVER.  executor.py sorts them by (start, end)
    (:1111-1112), but its cache transform (helpers/transforms.py:153-238, create_cache_hits) then moves
    the later spans of every cached trace earlier without re-sorting, so in this mode they arrive
    partly out of order, and the reference
      * builds the perfect-cut windows on those lists (bisect on unsorted data, V3:1115),
      * sorts out_span_partitions in place in TallySkipSpans (V3:968-971) — the no-deletion search
        (V3:1185) and BuildDistributions see sorted lists,
      * but searches WITH deletion (V3:1182) on deep copies taken before the sort (V3:1104-1105).
    Out spans are named by their position in the caller's list throughout."""
    in_start = [int(x) for x in in_start]
    in_end = [int(x) for x in in_end]
    out_start = [[int(x) for x in o] for o in out_start]
    out_end = [[int(x) for x in o] for o in out_end]
    n, E = len(in_start), len(out_start)

    sets = feasible_sets(in_start, in_end, out_start, out_end, preds)
    windows = windows_from_sets(in_end, sets)
    window_ends = {w[1] for w in windows}

    # TallySkipSpans sorts every partition by float(start), stable (V3:968-971)
    order = [sorted(range(len(o)), key=lambda j, o=o: float(o[j])) for o in out_start]
    sorted_start = [[o[j] for j in od] for o, od in zip(out_start, order)]
    sorted_end = [[o[j] for j in od] for o, od in zip(out_end, order)]
    wins, budgets, skip_count = tally_skip_spans(in_start, in_end, sorted_start, list(time_windows_before))
    win_starts = [w[0] for w in wins]
    skip_base = [np.concatenate([[0], np.cumsum(c)]).astype(int).tolist() for c in skip_count]
    fetches = [[0] * len(wins) for _ in range(E)]
    normalized = any(b > 0 for b in budgets)

    samples, large_delay = build_distribution_samples(in_start, in_end, sorted_start, sorted_end)
    tab = pair_params(samples, E, values_before)

    def primary(b, e):        # AlsoNonPrimaryAncestor, V1:294-303
        return not any(x != b and x != e and b in preds[x] and x in preds[e] for x in range(E))

    def cost(a, b, t1, t2):   # GetEpPairCost, V1:117-139
        mean, std = tab[a, b]
        if np.isnan(mean):
            raise ReferenceUndefined(f"no distribution for the pair ({a}, {b})")
        if std < 1.0e-12:
            std = 0.001
        return norm_pdf(t2 - t1, mean, std) if normalized else norm_logpdf(t2 - t1, mean, std)

    def fetch_skip(e, key):   # FetchSkipFromWindow, V3:820-842
        cands = [s for s in win_starts if s <= key]
        if not cands:
            raise ReferenceUndefined("no time window starts at or before the in-span")
        w = win_starts.index(max(cands))
        c = skip_count[e][w]
        if c <= 0:
            return None
        g = skip_base[e][w] + fetches[e][w] % c      # least-used first == round robin
        fetches[e][w] += 1
        return -2 - g

    def score(i, tup):        # ScoreAssignmentAsPerInvocationGraph, V1:259-361
        if all(c < 0 for c in tup):
            raise ReferenceUndefined("all-skip tuple")
        real = [e for e in range(E) if tup[e] >= 0]
        last = max(real, key=lambda e: out_end[e][tup[e]])
        total, num = 0.0, 0
        for e in range(E):
            if tup[e] < 0:
                continue
            cs, ce = out_start[e][tup[e]], out_end[e][tup[e]]
            for b in preds[e]:
                if not primary(b, e):
                    continue
                if tup[b] < 0:
                    if len(preds[b]) == 0:               # FindValidAncestor -> None
                        total += cost(0, 1 + e, in_start[i], cs)
                    else:
                        valid = [a for a in preds[b] if tup[a] >= 0]
                        if not valid:
                            raise ReferenceUndefined("ancestor chain of skip spans")
                        la = max(valid, key=lambda a: out_end[a][tup[a]])
                        total += cost(1 + la, 1 + e, out_start[la][tup[la]], cs)   # (sic) the ancestor's START
                    num += 1
                    continue
                total += cost(1 + b, 1 + e, out_end[b][tup[b]], cs)
                num += 1
            if len(preds[e]) == 0:
                total += cost(0, 1 + e, in_start[i], cs)
                num += 1
            if e == last:
                total += cost(1 + e, 0, ce, in_end[i])
                num += 1
        return total / num if normalized else total

    def topk(i, lists, count):
        """lists[e] = remaining original indices of ep e (ascending).  Returns (sorted entries, #leaves)."""
        heap, leaves = [], [0]
        s_in, e_in = in_start[i], in_end[i]

        def rec(level, tup):
            if level == E:
                leaves[0] += 1
                sc = score(i, tup)
                starts = [out_start[e][c] if c >= 0 else None for e, c in enumerate(tup)]
                heapq.heappush(heap, _Entry(sc, list(tup), starts))
                if len(heap) > K:
                    heapq.heappop(heap)
                return
            for x in lists[level]:
                s, en = out_start[level][x], out_end[level][x]
                if s_in > s or en > e_in:
                    continue
                if any(tup[b] >= 0 and out_end[b][tup[b]] > s for b in preds[level]):
                    continue
                rec(level + 1, tup + [x])
            code = fetch_skip(level, s_in)               # the None sentinel, V3:231-234, :316-320
            if code is not None:
                rec(level + 1, tup + [code])
        rec(0, [])
        heap.sort(reverse=True)
        return heap, leaves[0]

    remaining = [list(range(len(o))) for o in out_start]      # copies taken before the sort: caller's order
    full = [list(od) for od in order]                         # sorted in place
    assign = np.full((E, n), -1, np.int32)
    mis_rank = np.full(n, -1, np.int8)
    n_cand = np.zeros(n, np.int64)
    tk_score = np.full((n, K), np.nan)
    tk_idx = np.full((n, K, E), -1, np.int32)
    tk_cnt = np.zeros(n, np.int32)
    t2_score = np.full((n, K), np.nan)
    t2_idx = np.full((n, K, E), -1, np.int32)
    t2_cnt = np.zeros(n, np.int32)
    not_best = unassigned = 0
    batch = []
    for i in range(n):
        top, leaves = topk(i, remaining, True)
        n_cand[i] = leaves
        top2, _ = topk(i, full, False)
        for r, en in enumerate(top):
            tk_score[i, r], tk_idx[i, r] = en.score, en.tup
        tk_cnt[i] = len(top)
        for r, en in enumerate(top2):
            t2_score[i, r], t2_idx[i, r] = en.score, en.tup
        t2_cnt[i] = len(top2)
        batch.append((i, top))
        if i in window_ends:
            chosen = (mwis or exact_mwis)([[(en.score, en.tup) for en in t] for _, t in batch])
            for (ii, t), r in zip(batch, chosen):
                mis_rank[ii] = r
                if r < 0:
                    unassigned += 1
                    not_best += 1
                    continue
                if r != 0:
                    not_best += 1
                for e, c in enumerate(t[r].tup):
                    assign[e, ii] = c if c >= 0 else -2
                    if c >= 0:
                        remaining[e].remove(c)
            batch = []
    return dict(assign=assign, mis_rank=mis_rank, n_cand=n_cand, topk_score=tk_score, topk_idx=tk_idx,
                topk_cnt=tk_cnt, topk2_score=t2_score, topk2_idx=t2_idx, topk2_cnt=t2_cnt,
                not_best_count=not_best, cnt_unassigned=unassigned, windows=windows, time_windows=wins,
                skip_budget=budgets, skip_count=skip_count, pair_params=tab, large_delay=large_delay,
                normalized=normalized, samples=samples)


MWIS_FIXED_SCALE = 2.0 ** 42


def exact_mwis(cands):
    """BuildMISInstance V3:1252-1274 + an exact solver (the reference: gurobi_optimods.mwis, V3:1411).
    cands[k] = [(score, tuple)] ranks of in-span k.  Vertices of one in-span form a clique; two
    vertices of different in-spans conflict when they hold the same span (or the same skip span) at
    a tuple position (AssignmentIntersect, V3:1276-1281).

    Normalised scores are averages of densities (~1e-4), so the vertex weights 10000 + score differ
    in their last few bits only and a floating-point total depends on the order of the additions.
    The weights are therefore added EXACTLY: a double in [2^10, 2^14) is an integer multiple of
    2^-42, so weight * 2^42 is an integer and totals of a window (<= 31 vertices) fit 64 bits.  The
    optimum of the instance as the solver is given it is then well defined; of several optimal sets
    the first in depth-first order (in-spans ascending, ranks ascending, "unassigned" last) wins."""
    nw = len(cands)
    w = [[int((WEIGHT_OFFSET + s) * MWIS_FIXED_SCALE) for s, _ in c] for c in cands]

    def conflict(a, ra, b, rb):
        return any(x == y for x, y in zip(cands[a][ra][1], cands[b][rb][1]))

    # connected components of the in-span conflict graph are independent instances; solving them one
    # by one (members ascending) gives the same set as one depth-first search over the whole window
    adj = [[any(conflict(a, ra, b, rb) for ra in range(len(cands[a])) for rb in range(len(cands[b])))
            if a != b else False for b in range(nw)] for a in range(nw)]
    chosen = [-1] * nw
    seen = [False] * nw
    for seed in range(nw):
        if seen[seed]:
            continue
        comp, stack = [], [seed]
        seen[seed] = True
        while stack:
            k = stack.pop()
            comp.append(k)
            for j in range(nw):
                if adj[k][j] and not seen[j]:
                    seen[j] = True
                    stack.append(j)
        comp.sort()
        m = len(comp)
        ub = [0] * (m + 1)
        for l in range(m - 1, -1, -1):
            ub[l] = ub[l + 1] + max([0] + w[comp[l]])
        best = [-1, [-1] * m]
        cur = [-1] * m

        def rec(l, tot):
            if l == m:
                if tot > best[0]:
                    best[0], best[1] = tot, list(cur)
                return
            if tot + ub[l] <= best[0]:
                return
            k = comp[l]
            for r in range(len(cands[k])):
                if not w[k][r] > 0:
                    continue
                if any(cur[a] >= 0 and conflict(comp[a], cur[a], k, r) for a in range(l)):
                    continue
                cur[l] = r
                rec(l + 1, tot + w[k][r])
            cur[l] = -1
            rec(l + 1, tot)
        rec(0, 0)
        for l, k in enumerate(comp):
            chosen[k] = best[1][l]
    return chosen
