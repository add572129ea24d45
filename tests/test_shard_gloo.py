"""N>1 host logic on CPU: two gloo ranks shard a service list, reduce timings with MAX, gather
variable-length assignment arrays.  No GPU."""
import os
import socket

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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.service_range(25, rank, world)
    blocks = synth.hotel_stream(12, 50, seed=shard.shard_seed(10, rank))
    hb = build_batch_from_blocks(blocks)
    spans = synth.span_count(blocks)
    slow = shard.max_over_ranks(10.0 + rank)
    total = shard.sum_over_ranks(float(spans))
    fake_assign = torch.arange(5 + 3 * rank, dtype=torch.int32) + 100 * rank
    parts = shard.gather_assignments(fake_assign)
    q.put((rank, lo, hi, int(hb.in_start[0]), spans, slow, total, [p.tolist() for p in parts]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, first0, sp0, slow0, tot0, parts0), (r1, lo1, hi1, first1, sp1, slow1, tot1, parts1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 13, 13, 25)               # disjoint, covering, balanced
    assert first0 != first1                                        # different shards of the stream
    assert slow0 == slow1 == 11.0                                  # MAX over ranks
    assert tot0 == tot1 == sp0 + sp1
    assert parts0 == parts1 == [list(range(5)), [100 + i for i in range(8)]]


def test_service_range_partitions():
    for n in (1, 7, 8, 100):
        for w in (1, 2, 3, 8):
            got = [shard.service_range(n, r, w) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1
