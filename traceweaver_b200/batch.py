"""Host-side problem/batch descriptors: the SoA layout the engine consumes.

A *problem* is what the reference hands to `TraceWeaverV3.FindAssignments`
(executor.py:1172-1175): the spans arriving at one service (one incoming endpoint) and the spans
it sends to each outgoing endpoint, plus the invocation DAG over the outgoing endpoints.  Here it
is index-only: int64 microsecond start/end arrays sorted by (start, end) (executor.py:1111-1112),
endpoints in topological order (traceweaver_v1.py:37-39), DAG as predecessor lists in
`in_edges` order (the order the reference sums likelihood terms in, traceweaver_v1.py:322).

`build_batch` concatenates problems into the arrays of `tw_batch` (include/traceweaver_b200.h).
Pure numpy; no device code here.
"""
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _abi


@dataclass
class Problem:
    in_start: np.ndarray                 # int64 [n_in]
    in_end: np.ndarray                   # int64 [n_in]
    out_start: List[np.ndarray]          # per ep (topological order): int64 [n_out_e]
    out_end: List[np.ndarray]
    preds: List[List[int]]               # per ep: predecessor positions, in_edges order
    name: str = ""

    @property
    def E(self):
        return len(self.out_start)

    @property
    def n_in(self):
        return int(self.in_start.shape[0])

    def is_primary(self, b, e):
        """Edge b->e is NON-primary iff a 2-hop path b->x->e exists
        (AlsoNonPrimaryAncestor, traceweaver_v1.py:294-303: all_simple_paths(cutoff=2))."""
        for x in range(self.E):
            if x != b and x != e and b in self.preds[x] and x in self.preds[e]:
                return False
        return True

    def terms(self):
        """[(ep, src)] in the reference's summation order (traceweaver_v1.py:316-357):
        per ep: primary in-edges in in_edges order | ROOT if no in-edges; then LAST."""
        out = []
        for e in range(self.E):
            for b in self.preds[e]:
                if self.is_primary(b, e):
                    out.append((e, b))
            if len(self.preds[e]) == 0:
                out.append((e, _abi.TW_TERM_ROOT))
            out.append((e, _abi.TW_TERM_LAST))
        return out

    def validate(self):
        E = self.E
        if not (1 <= E <= _abi.TW_MAX_E):
            raise ValueError(f"{self.name}: E={E} outside [1, {_abi.TW_MAX_E}]")
        if self.n_in < 2:
            # the reference builds no window for a single in-span and then fails on max([])
            # (traceweaver_v3.py:1056-1076, :1119)
            raise ValueError(f"{self.name}: need at least 2 incoming spans")
        for e in range(E):
            for b in self.preds[e]:
                if not (0 <= b < e):
                    raise ValueError(f"{self.name}: preds must reference earlier topological positions")
        if len(self.out_end) != E or len(self.preds) != E:
            raise ValueError(f"{self.name}: out_start / out_end / preds must have one entry per ep")
        if self.in_end.shape != self.in_start.shape:
            raise ValueError(f"{self.name}: in_start and in_end differ in length")
        if np.any(self.in_end < self.in_start):
            raise ValueError(f"{self.name}: an in-span ends before it starts")
        key = self.in_start.astype(np.int64)
        if np.any(np.diff(key) < 0):
            raise ValueError(f"{self.name}: in-spans not sorted by start")
        for e in range(E):
            if self.out_end[e].shape != self.out_start[e].shape:
                raise ValueError(f"{self.name}: out_start and out_end of ep {e} differ in length")
            if np.any(self.out_end[e] < self.out_start[e]):
                raise ValueError(f"{self.name}: an out-span of ep {e} ends before it starts")
            if np.any(np.diff(self.out_start[e]) < 0):
                raise ValueError(f"{self.name}: out-spans of ep {e} not sorted by start")

    def in_accelerated_regime(self):
        """True iff every ep has as many outgoing spans as there are incoming ones (no skip budget,
        traceweaver_v3.py:1138-1158): the regime the engine solves."""
        return all(len(o) == self.n_in for o in self.out_start)


@dataclass
class HostBatch:
    problems: Sequence[Problem]
    arrays: dict = field(default_factory=dict)

    def __getattr__(self, k):
        try:
            return self.__dict__["arrays"][k]
        except KeyError:
            raise AttributeError(k)

    @property
    def n_problems(self):
        return len(self.problems)

    def no_skip(self):
        """True iff every ep of every problem has n_out == n_in (the regime with two passes,
        traceweaver_v3.py:1138-1158)."""
        a = self.arrays
        n_in = np.diff(a["prob_in_off"])
        n_out = np.diff(a["ep_out_off"])
        ep_prob = np.repeat(np.arange(self.n_problems), np.diff(a["prob_ep_off"]))
        return bool(np.all(n_out == n_in[ep_prob]))

    def slice(self, lo: int, hi: int) -> "HostBatch":
        """Sub-batch of problems [lo, hi): offset tables rebased, span arrays are views (no copy).
        Services are independent problems, so solving the slices separately gives the same result."""
        a = self.arrays
        e0, e1 = int(a["prob_ep_off"][lo]), int(a["prob_ep_off"][hi])
        t0, t1 = int(a["ep_term_off"][e0]), int(a["ep_term_off"][e1])
        i0, i1 = int(a["prob_in_off"][lo]), int(a["prob_in_off"][hi])
        o0, o1 = int(a["ep_out_off"][e0]), int(a["ep_out_off"][e1])

        def reb(name, x0, x1):
            v = a[name][x0:x1 + 1]
            return np.ascontiguousarray(v - v[0])

        arrays = dict(
            prob_in_off=reb("prob_in_off", lo, hi), prob_ep_off=reb("prob_ep_off", lo, hi),
            prob_tuple_off=reb("prob_tuple_off", lo, hi), ep_out_off=reb("ep_out_off", e0, e1),
            ep_term_off=reb("ep_term_off", e0, e1), ep_pred_mask=a["ep_pred_mask"][e0:e1],
            term_src=a["term_src"][t0:t1], in_start=a["in_start"][i0:i1], in_end=a["in_end"][i0:i1],
            out_start=a["out_start"][o0:o1], out_end=a["out_end"][o0:o1],
            prob_gauss_off=reb("prob_gauss_off", lo, hi), term_sample_off=reb("term_sample_off", t0, t1))
        probs = self.problems[lo:hi] if isinstance(self.problems, list) else _ProblemCount(hi - lo)
        return HostBatch(problems=probs, arrays=arrays)


def build_batch(problems: Sequence[Problem], validate=True) -> HostBatch:
    P = len(problems)
    if P == 0:
        raise ValueError("empty batch")
    if validate:
        for p in problems:
            p.validate()
    prob_in_off = np.zeros(P + 1, np.int64)
    prob_ep_off = np.zeros(P + 1, np.int32)
    prob_tuple_off = np.zeros(P + 1, np.int64)
    ep_out_off = [0]
    ep_term_off = [0]
    ep_pred_mask = []
    term_src = []
    for i, p in enumerate(problems):
        prob_in_off[i + 1] = prob_in_off[i] + p.n_in
        prob_ep_off[i + 1] = prob_ep_off[i] + p.E
        prob_tuple_off[i + 1] = prob_tuple_off[i] + p.n_in * p.E
        terms = p.terms()
        for e in range(p.E):
            ep_out_off.append(ep_out_off[-1] + int(p.out_start[e].shape[0]))
            mask = 0
            for b in p.preds[e]:
                mask |= 1 << b
            ep_pred_mask.append(mask)
            mine = [src for (ee, src) in terms if ee == e]
            term_src.extend(mine)
            ep_term_off.append(ep_term_off[-1] + len(mine))
    arrays = dict(
        prob_in_off=prob_in_off, prob_ep_off=prob_ep_off, prob_tuple_off=prob_tuple_off,
        ep_out_off=np.asarray(ep_out_off, np.int64), ep_term_off=np.asarray(ep_term_off, np.int32),
        ep_pred_mask=np.asarray(ep_pred_mask, np.uint32), term_src=np.asarray(term_src, np.int8),
        in_start=np.ascontiguousarray(np.concatenate([p.in_start for p in problems]), np.int64),
        in_end=np.ascontiguousarray(np.concatenate([p.in_end for p in problems]), np.int64),
        out_start=np.ascontiguousarray(np.concatenate([s for p in problems for s in p.out_start]), np.int64),
        out_end=np.ascontiguousarray(np.concatenate([s for p in problems for s in p.out_end]), np.int64),
    )
    n_batches = (np.diff(prob_in_off) + _abi.TW_PARAM_BATCH - 1) // _abi.TW_PARAM_BATCH
    n_terms = np.diff(np.asarray(ep_term_off, np.int64)[prob_ep_off])
    arrays["prob_gauss_off"] = np.concatenate([[0], np.cumsum(n_batches * n_terms)]).astype(np.int64)
    # sample capacity of a term = n_in of its problem (tw_delays)
    term_prob = np.repeat(np.arange(P), n_terms)
    arrays["term_sample_off"] = np.concatenate(
        [[0], np.cumsum(np.diff(prob_in_off)[term_prob])]).astype(np.int64)
    return HostBatch(problems=list(problems), arrays=arrays)


def batch_struct(hb: HostBatch, ptr):
    """Fill a TwBatch with pointers produced by `ptr(name)` (host or device)."""
    a = hb.arrays
    s = _abi.TwBatch()
    s.n_problems = hb.n_problems
    s.n_ep_total = int(a["prob_ep_off"][-1])
    s.n_term_total = int(a["ep_term_off"][-1])
    s.n_in_total = int(a["prob_in_off"][-1])
    s.n_out_total = int(a["ep_out_off"][-1])
    for name in ("prob_in_off", "prob_ep_off", "prob_tuple_off", "ep_out_off", "ep_term_off",
                 "ep_pred_mask", "term_src", "in_start", "in_end", "out_start", "out_end"):
        setattr(s, name, ptr(name))
    return s


@dataclass
class ServiceBlock:
    """S services with identical shape (n_in, E, DAG): the vectorised form used by the synthetic
    generators.  Arrays are [S, n]; out lists are per ep in topological order."""
    in_start: np.ndarray
    in_end: np.ndarray
    out_start: List[np.ndarray]
    out_end: List[np.ndarray]
    preds: List[List[int]]
    truth: np.ndarray = None            # [E, S, n] index of the true child in each ep list
    name: str = ""

    def problem(self, s) -> Problem:
        return Problem(in_start=self.in_start[s], in_end=self.in_end[s],
                       out_start=[o[s] for o in self.out_start], out_end=[o[s] for o in self.out_end],
                       preds=self.preds, name=f"{self.name}[{s}]")


def build_batch_from_blocks(blocks: Sequence[ServiceBlock]) -> HostBatch:
    """Same arrays as build_batch([...problems of every block...]) without per-problem Python work."""
    prob_in, prob_ep, prob_tuple = [0], [0], [0]
    ep_out, ep_term, ep_pred, term_src = [0], [0], [], []
    ins, ine, outs, oute = [], [], [], []
    for blk in blocks:
        S, n = blk.in_start.shape
        E = len(blk.out_start)
        tmpl = Problem(in_start=blk.in_start[0], in_end=blk.in_end[0], out_start=[o[0] for o in blk.out_start],
                       out_end=[o[0] for o in blk.out_end], preds=blk.preds, name=blk.name)
        tmpl.validate()
        terms = tmpl.terms()
        per_ep_terms = [[src for (ee, src) in terms if ee == e] for e in range(E)]
        masks = [sum(1 << b for b in blk.preds[e]) for e in range(E)]
        base_in, base_ep, base_tuple = prob_in[-1], prob_ep[-1], prob_tuple[-1]
        prob_in.extend(base_in + n * np.arange(1, S + 1))
        prob_ep.extend(base_ep + E * np.arange(1, S + 1))
        prob_tuple.extend(base_tuple + n * E * np.arange(1, S + 1))
        base_out, base_term = ep_out[-1], ep_term[-1]
        ep_out.extend(base_out + n * np.arange(1, S * E + 1))
        cum = np.cumsum([len(t) for t in per_ep_terms])
        nt = int(cum[-1])
        ep_term.extend((base_term + nt * np.arange(S)[:, None] + cum[None, :]).reshape(-1))
        ep_pred.extend(masks * S)
        term_src.extend([src for t in per_ep_terms for src in t] * S)
        ins.append(blk.in_start.reshape(-1))
        ine.append(blk.in_end.reshape(-1))
        # per problem: ep 0 list, ep 1 list, ...  -> stack [S, E, n]
        outs.append(np.stack(blk.out_start, axis=1).reshape(-1))
        oute.append(np.stack(blk.out_end, axis=1).reshape(-1))
    arrays = dict(
        prob_in_off=np.asarray(prob_in, np.int64), prob_ep_off=np.asarray(prob_ep, np.int32),
        prob_tuple_off=np.asarray(prob_tuple, np.int64), ep_out_off=np.asarray(ep_out, np.int64),
        ep_term_off=np.asarray(ep_term, np.int32), ep_pred_mask=np.asarray(ep_pred, np.uint32),
        term_src=np.asarray(term_src, np.int8),
        in_start=np.ascontiguousarray(np.concatenate(ins), np.int64),
        in_end=np.ascontiguousarray(np.concatenate(ine), np.int64),
        out_start=np.ascontiguousarray(np.concatenate(outs), np.int64),
        out_end=np.ascontiguousarray(np.concatenate(oute), np.int64))
    P = len(prob_in) - 1
    n_in = np.diff(arrays["prob_in_off"])
    n_batches = (n_in + _abi.TW_PARAM_BATCH - 1) // _abi.TW_PARAM_BATCH
    n_terms = np.diff(arrays["ep_term_off"].astype(np.int64)[arrays["prob_ep_off"]])
    arrays["prob_gauss_off"] = np.concatenate([[0], np.cumsum(n_batches * n_terms)]).astype(np.int64)
    term_prob = np.repeat(np.arange(P), n_terms)
    arrays["term_sample_off"] = np.concatenate([[0], np.cumsum(n_in[term_prob])]).astype(np.int64)
    hb = HostBatch(problems=_ProblemCount(P), arrays=arrays)
    return hb


class _ProblemCount:
    """len()-only stand-in for the problem list of a block-built batch."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n
