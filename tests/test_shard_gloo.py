"""N>1 host logic on CPU (gloo, world_size 2): ONE service list is partitioned by span count, each
rank generates and solves only its slice (here with the CPU oracle standing in for the engine), and
the path's single collective — the all-gather of the assignment arrays — leaves every rank with the
assignments of the whole list, equal to a single-process solve.  No GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from traceweaver_b200 import shard, synth
from traceweaver_b200.batch import build_batch_from_blocks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


N_SERVICES, N_IN, SEED = 26, 60, 10


def _specs():
    return shard.stream_spec("hotel", N_SERVICES, N_IN, SEED, block_services=3)


def _solve(blocks):
    from oracle import tw_oracle
    hb = build_batch_from_blocks(blocks)
    return tw_oracle.find_assignments(hb, SEED, threads=1, want_topk=False)["assign"]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    specs = _specs()
    bounds = shard.partition_by_spans(shard.spec_span_counts(specs), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    blocks = shard.generate_slice(specs, lo, hi)
    assign = torch.from_numpy(_solve(blocks))
    gather = shard.AssignGather(shard.spec_tuple_counts(specs), bounds, torch.device("cpu"))
    buf = gather(assign, rank)
    whole = torch.cat(gather.shards(buf)).numpy()
    slow = shard.max_over_ranks(10.0 + rank)
    total = shard.sum_over_ranks(float(synth.span_count(blocks)))
    q.put((rank, lo, hi, synth.span_count(blocks), slow, total, whole.tolist(), gather.bytes_received_per_rank))
    dist.destroy_process_group()


def test_two_ranks_partition_solve_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, sp0, slow0, tot0, whole0, rx0), (r1, lo1, hi1, sp1, slow1, tot1, whole1, rx1) = out
    specs = _specs()
    counts = shard.spec_span_counts(specs)
    assert lo0 == 0 and hi0 == lo1 and hi1 == N_SERVICES                     # disjoint, covering, contiguous
    assert abs(sp0 - sp1) <= counts.max()                                    # balanced by SPAN count
    assert sp0 == counts[lo0:hi0].sum() and sp1 == counts[lo1:hi1].sum()
    assert slow0 == slow1 == 11.0                                            # MAX over ranks
    assert tot0 == tot1 == counts.sum()
    # every rank ends up with the assignments of the WHOLE list == one process solving it alone
    single = _solve(shard.generate_slice(specs, 0, N_SERVICES))
    assert whole0 == whole1 == single.tolist()
    assert rx0 == rx1 > 0


def test_partition_by_spans_properties():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 100, 1000):
        counts = rng.integers(1, 5000, size=n)
        for w in (1, 2, 3, 8):
            b = shard.partition_by_spans(counts, w)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
            loads = [counts[b[r]:b[r + 1]].sum() for r in range(w)]
            assert sum(loads) == counts.sum()
            if n >= 8 * w:
                assert max(loads) - counts.sum() / w <= counts.max()       # within one service of ideal
            assert shard.service_range(counts, 0, w) == (0, int(b[1]))


def test_slice_generation_is_cut_independent():
    """A rank's slice must not depend on where the other ranks' boundaries fall."""
    specs = _specs()
    whole = build_batch_from_blocks(shard.generate_slice(specs, 0, N_SERVICES))
    for lo, hi in ((0, 5), (5, 19), (19, 26), (7, 8)):
        part = build_batch_from_blocks(shard.generate_slice(specs, lo, hi))
        i0, i1 = int(whole.prob_in_off[lo]), int(whole.prob_in_off[hi])
        assert np.array_equal(part.in_start, whole.in_start[i0:i1])
        e0, e1 = int(whole.prob_ep_off[lo]), int(whole.prob_ep_off[hi])
        o0, o1 = int(whole.ep_out_off[e0]), int(whole.ep_out_off[e1])
        assert np.array_equal(part.out_end, whole.out_end[o0:o1])
        assert np.array_equal(part.ep_pred_mask, whole.ep_pred_mask[e0:e1])
