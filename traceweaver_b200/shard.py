"""Multi-GPU host logic: the span stream shards by service (services are independent problems,
executor.py:1080), one process per GPU, no collective on the data path.  torch.distributed is used
for rendezvous, barriers and reducing timings/counters only (NCCL on GPUs, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def service_range(n_services_total: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of the service list owned by `rank`."""
    base, rem = divmod(n_services_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seed(seed: int, rank: int) -> int:
    """Seed of rank's synthetic shard (disjoint streams per rank)."""
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
    """Optional consumer-side gather of per-rank assignment arrays (variable length) to every rank."""
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
