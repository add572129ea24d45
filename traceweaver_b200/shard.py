"""Multi-GPU host logic: ONE service list is partitioned across the ranks by span
count — services are independent problems (executor.py:1080), so a rank solves its contiguous slice
with no halo — and every pass over the list ends with a single all-gather of the per-service
assignment arrays, after which every rank holds the assignments of the whole list (what the
reference's consumers, helpers/utils.py AccuracyEndToEnd / ConstructEndToEndTraces, need: they join
the per-service maps across services).  One process per GPU; torch.distributed for the rendezvous,
the collective (NCCL on GPUs, gloo in the CPU tests) and for reducing timings.
"""
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def partition_by_spans(span_counts: Sequence[int], world: int) -> np.ndarray:
    """Boundaries [world+1] of contiguous service ranges with (nearly) equal span counts: rank r owns
    services [b[r], b[r+1]).  Greedy on the cumulative count: a service goes to the rank whose ideal
    share its midpoint falls into, so no rank exceeds the ideal by more than half a service."""
    c = np.asarray(span_counts, np.int64)
    cum = np.concatenate([[0], np.cumsum(c)])
    total = int(cum[-1])
    mid = (cum[:-1] + cum[1:]) / 2.0
    owner = np.minimum((mid * world / max(total, 1)).astype(np.int64), world - 1)
    bounds = np.searchsorted(owner, np.arange(world + 1), side="left")
    bounds[0], bounds[-1] = 0, len(c)
    return bounds.astype(np.int64)


def service_range(span_counts, rank: int, world: int):
    """[lo, hi) of the services owned by `rank` (partition by span count).  An int argument means
    that many services of equal size."""
    if np.isscalar(span_counts):
        span_counts = np.ones(int(span_counts), np.int64)
    b = partition_by_spans(span_counts, world)
    return int(b[rank]), int(b[rank + 1])


@dataclass
class BlockSpec:
    """A run of `services` identically shaped synthetic services (synth.make_block arguments)."""
    shape: str
    services: int
    n_in: int
    load: float
    seed: int
    quantum_us: int = 1

    def spans_per_service(self):
        from . import synth
        return self.n_in * (1 + len(synth.SHAPES[self.shape]["eps"]))

    def tuples_per_service(self):
        from . import synth
        return self.n_in * len(synth.SHAPES[self.shape]["eps"])


def stream_spec(workload: str, n_services: int, n_in: int, seed: int, block_services: int = 32) -> List[BlockSpec]:
    """The global service list of a synthetic workload as small seeded blocks, so that a rank can
    generate exactly its slice: block k is reproducible from (shape, load, seed + k) alone.  Blocks are
    short (one (shape, load) cycle = 12 blocks = 384 services for the hotel stream) so that every
    contiguous slice of a few thousand services holds the same mix of shapes and load levels: with long
    blocks the ranks of a partitioned list got different mixes and the slowest rank set the step time."""
    loads = (25, 50, 75, 100, 125, 150)
    if workload == "hotel":
        cycle = [(s, l, 1) for l in loads for s in ("hotel_frontend", "hotel_search")]
    elif workload == "media":
        cycle = [(s, l, 1) for l in loads for s in ("media_nginx_cal", "media_movie_id", "media_leaf", "media_leaf",
                                                    "media_leaf", "media_leaf")]
    elif workload == "alibaba":
        rng = np.random.default_rng(seed)
        cycle = []
        for cf in (1, 200, 1000, 4000, 10000, 15000):
            for s in ("ali_leaf", "ali_chain2", "ali_par3", "ali_chain4", "ali_leaf", "ali_chain2"):
                replicas = 2 ** int(rng.integers(0, 13))
                cycle.append((s, min(float(max(1, int(np.ceil(cf / replicas)))), synth_max_load()), 1000))
    else:
        raise ValueError(workload)
    per = max(1, min(block_services, n_services // len(cycle)))
    specs, left, k = [], n_services, 0
    while left > 0:
        shape, load, q = cycle[k % len(cycle)]
        s = min(per, left)
        specs.append(BlockSpec(shape, s, n_in, load, seed + k, q))
        left -= s
        k += 1
    return specs


def synth_max_load() -> float:
    """Load cap of the alibaba-shaped generator (synth.alibaba_stream explains the cap)."""
    from . import synth
    return synth.ALIBABA_MAX_LOAD


def spec_span_counts(specs: Sequence[BlockSpec]) -> np.ndarray:
    return np.concatenate([np.full(b.services, b.spans_per_service(), np.int64) for b in specs])


def spec_tuple_counts(specs: Sequence[BlockSpec]) -> np.ndarray:
    return np.concatenate([np.full(b.services, b.tuples_per_service(), np.int64) for b in specs])


def generate_slice(specs: Sequence[BlockSpec], lo: int, hi: int):
    """ServiceBlocks of services [lo, hi) of the list `specs` describes (a block that straddles a
    boundary is generated whole and cut: its services do not depend on the cut)."""
    from . import synth
    from .batch import ServiceBlock
    blocks, first = [], 0
    for b in specs:
        a, z = max(lo, first), min(hi, first + b.services)
        if a < z:
            blk = synth.make_block(b.shape, b.services, b.n_in, b.load, b.seed, quantum_us=b.quantum_us)
            if z - a < b.services:
                r = slice(a - first, z - first)
                blk = ServiceBlock(in_start=np.ascontiguousarray(blk.in_start[r]), in_end=np.ascontiguousarray(blk.in_end[r]),
                                   out_start=[np.ascontiguousarray(o[r]) for o in blk.out_start],
                                   out_end=[np.ascontiguousarray(o[r]) for o in blk.out_end], preds=blk.preds,
                                   truth=np.ascontiguousarray(blk.truth[:, r]), name=blk.name)
            blocks.append(blk)
        first += b.services
    return blocks


class AssignGather:
    """The one collective of the data path: all-gather of the ranks' `assign` arrays (int32, one
    entry per (service, ep, in-span)).  Shard sizes follow from the partition, which every rank
    knows, so no size exchange happens at run time; shards are padded to the largest one because
    all_gather_into_tensor wants equal contributions."""

    def __init__(self, tuple_counts: Sequence[int], bounds: Sequence[int], device):
        cum = np.concatenate([[0], np.cumsum(np.asarray(tuple_counts, np.int64))])
        self.sizes = [int(cum[bounds[r + 1]] - cum[bounds[r]]) for r in range(len(bounds) - 1)]
        self.world = len(self.sizes)
        self.pad = max(self.sizes) if self.sizes else 0
        self.device = device
        self.buf = torch.empty(self.world * self.pad, dtype=torch.int32, device=device)
        self.mine = torch.full((self.pad,), -9, dtype=torch.int32, device=device)

    def __call__(self, assign: torch.Tensor, rank: int) -> torch.Tensor:
        """Returns the [world, pad] buffer; shard r is buf[r, :sizes[r]]."""
        assert assign.numel() == self.sizes[rank]
        if self.world == 1 or not dist.is_initialized():
            self.buf[: assign.numel()].copy_(assign)
            return self.buf.view(1, -1)
        self.mine[: assign.numel()].copy_(assign)
        dist.all_gather_into_tensor(self.buf, self.mine)
        return self.buf.view(self.world, self.pad)

    def shards(self, buf: torch.Tensor):
        return [buf[r, : self.sizes[r]] for r in range(self.world)]

    @property
    def bytes_received_per_rank(self):
        return 4 * self.pad * (self.world - 1)


def shard_seed(seed: int, rank: int) -> int:
    """Seed of an independent per-rank stream (replica mode; the partitioned stream does not use it)."""
    return seed + 1000 * rank


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's time: the only timing that is valid for a multi-GPU step."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_assignments(assign: torch.Tensor):
    """Gather of per-rank assignment arrays whose lengths are NOT known in advance (one size
    exchange, then a padded all-gather).  The bench path uses AssignGather instead."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [assign]
    world = dist.get_world_size()
    n = torch.tensor([assign.numel()], dtype=torch.int64, device=assign.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    pad = torch.full((m,), -9, dtype=assign.dtype, device=assign.device)
    pad[: assign.numel()] = assign
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return [b[: int(s.item())] for b, s in zip(bufs, sizes)]
