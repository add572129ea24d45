"""The accuracy helpers of traceweaver_b200.loader (index-array restatement of helpers/utils.py:34-145)
against the numbers executor.py printed when the goldens were minted: per-service accuracy, top-K
accuracy and the two end-to-end accuracies, all from the committed fixtures."""
import glob
import json
import os
import re

import numpy as np
import pytest

from golden_util import Golden, GOLDEN_DIR

DATASETS = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*_load*.json")))


def _printed(lines, pattern):
    out = {}
    for ln in lines:
        m = re.match(pattern, ln)
        if m:
            out[m.group(1)] = float(m.group(2))
    return out


@pytest.mark.parametrize("path", DATASETS, ids=[os.path.basename(p)[:-5] for p in DATASETS])
def test_accuracies_equal_the_reference_printout(path):
    from traceweaver_b200.loader import (ServiceProblem, accuracy, topk_accuracy, end_to_end_accuracy,
                                         end_to_end_topk_accuracy)
    meta = json.load(open(path))
    lines = meta["printed_accuracy"]
    per = _printed(lines, r"Accuracy for service (.+): ([0-9.]+)%")
    per_k = _printed(lines, r"Top K accuracy for service (.+): ([0-9.]+)%")
    e2e = _printed(lines, r"End-to-end accuracy for method (MaxScoreBatchSubsetWithSkips): ([0-9.]+)%")
    e2e_k = _printed(lines, r"End-to-end top K accuracy for method (MaxScoreBatchSubsetWithSkips): ([0-9.]+)%")
    order = list(meta["find_assignments_seconds"].keys())              # the order the services were solved in
    gs = {n: Golden(os.path.join(GOLDEN_DIR, f"{meta['dataset']}__{n}.npz")) for n in order}
    for name, g in gs.items():
        z = g.z
        sp = ServiceProblem(name=name, in_ep="", out_eps_given=[], out_eps=g.topo, problem=None, in_ids=[],
                            out_ids=[], truth=z["truth"])
        assert round(100 * accuracy(sp, z["assign"]), 3) == per[name]
        assert round(100 * topk_accuracy(z["truth"], z["topk_final"], z["topk_final_cnt"]), 3) == per_k[name]
    tr = [list(gs[n].z["in_trace"]) for n in order]
    tru = [gs[n].z["truth"] for n in order]
    got = end_to_end_accuracy(tr, tru, [gs[n].z["assign"] for n in order])
    assert round(100 * got, 3) == e2e["MaxScoreBatchSubsetWithSkips"]
    got_k = end_to_end_topk_accuracy(tr, tru, [gs[n].z["topk_final"] for n in order],
                                     [gs[n].z["topk_final_cnt"] for n in order])
    assert round(100 * got_k, 3) == e2e_k["MaxScoreBatchSubsetWithSkips"]
