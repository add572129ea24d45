"""Times the data path's collective alone (all_gather_into_tensor of int32 shards) on N ranks."""
import os, sys, time
import torch, torch.distributed as dist
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_480_000
mine = torch.full((n,), rank, dtype=torch.int32, device=dev)
buf = torch.empty(world * n, dtype=torch.int32, device=dev)
for _ in range(3):
    dist.all_gather_into_tensor(buf, mine)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    dist.all_gather_into_tensor(buf, mine)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
if rank == 0:
    print(f"all_gather {4*n/1e6:.1f} MB per rank, world {world}: {ms:.3f} ms  -> {4*n*(world-1)/ms/1e6:.1f} GB/s received per rank", flush=True)
    print(torch.cuda.get_device_name(lr))
dist.destroy_process_group()
