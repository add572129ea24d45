#!/usr/bin/env python
"""Synthetic traces in the JSON layout the reference's Alibaba ETL emits
(src/trace_reconstructor/ports/python/alibaba-analysis/real-parser.py:308-359: one server record per
rpc — processID = callee, spanID = the dotted rpc id, CHILD_OF reference to the parent rpc id — plus,
for every rpc but the root, a client twin with the same spanID and processID = caller; `caller`,
`callee`, `requestType` fields; no `processes` table).  The real trace is not shipped (Git LFS pointer),
so the `--fix 5` layout of the loader (executor.py:377-399: `.client` ids, self-loop renaming) is pinned
on these: a small call graph with a nested call, two sequential callees and a SELF LOOP (S2 -> S2);
five of the 160 traces violate the parent/child time containment and are dropped by the reference.

    root 0: client -> S0
      0.1: S0 -> S1          0.1.1: S1 -> S3
      0.2: S0 -> S2          0.2.1: S2 -> S2 (self loop)     0.2.1.1: S2 -> S4

Usage: python tests/golden/make_alibaba_traces.py   (writes tests/golden/alibaba_synth/*.json, seeded)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "alibaba_synth")
N_TRACES = 160
SEED = 5


def record(trace_id, rpc, caller, callee, start, dur, kind, parent):
    refs = [] if parent is None else [{"refType": "CHILD_OF", "traceID": trace_id, "spanID": parent}]
    return {"traceID": trace_id, "startTime": int(start), "spanID": rpc, "caller": caller, "requestType": "rpc",
            "callee": callee, "interface": "iface", "duration": int(dur),
            "tags": [{"key": "span.kind", "value": kind}], "references": refs,
            "processID": callee if kind == "server" else caller}


def main():
    rng = np.random.default_rng(SEED)
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        if f.endswith(".json"):
            os.remove(os.path.join(OUT, f))
    t = 1_655_000_000_000_000
    for k in range(N_TRACES):
        t += int(rng.integers(1500, 6000))                        # arrivals: overlapping requests
        tid = f"trace{k:04d}"

        def ln(mu, sigma):
            return int(max(50, rng.lognormal(np.log(mu), sigma)))
        spans = []
        # leaf work first, parents enclose their children (the loader checks containment, executor.py:425-438)
        s0 = t
        a1 = s0 + ln(300, 0.4)                                   # 0.1: S0 -> S1
        a11 = a1 + ln(200, 0.4)                                  # 0.1.1: S1 -> S3
        d11 = ln(900, 0.5)
        d1 = (a11 - a1) + d11 + ln(250, 0.4)
        a2 = a1 + d1 + ln(350, 0.4)                              # 0.2: S0 -> S2, after 0.1 returned
        a21 = a2 + ln(200, 0.4)                                  # 0.2.1: S2 -> S2
        a211 = a21 + ln(150, 0.4)                                # 0.2.1.1: S2 -> S4
        d211 = ln(700, 0.5)
        d21 = (a211 - a21) + d211 + ln(200, 0.4)
        d2 = (a21 - a2) + d21 + ln(250, 0.4)
        d0 = (a2 - s0) + d2 + ln(300, 0.4)
        if k % 37 == 11:
            d211 = d21 + 500          # the innermost call outlives its caller: the reference drops such a trace
                                      # (check_time_constraints, executor.py:425-441), and so must the loader
        calls = [("0", None, "client", "S0", s0, d0), ("0.1", "0", "S0", "S1", a1, d1),
                 ("0.1.1", "0.1", "S1", "S3", a11, d11), ("0.2", "0", "S0", "S2", a2, d2),
                 ("0.2.1", "0.2", "S2", "S2", a21, d21), ("0.2.1.1", "0.2.1", "S2", "S4", a211, d211)]
        for rpc, parent, caller, callee, st, du in calls:
            spans.append(record(tid, rpc, caller, callee, st, du, "server", parent))
            if parent is not None:
                spans.append(record(tid, rpc, caller, callee, st, du, "client", parent))
        with open(os.path.join(OUT, tid + ".json"), "w") as f:
            json.dump({"data": [{"traceID": tid, "spans": spans}]}, f)
    print("wrote", N_TRACES, "traces to", OUT)


if __name__ == "__main__":
    main()
