"""Times the training step's collective in isolation: all-reduce of the flat gradient bucket (debug aid)."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (1, 124008, 11_100_000):
    x = torch.ones(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        print(f"all_reduce {n} fp32: {e0.elapsed_time(e1) / 20:.3f} ms device, {(time.time() - t0) / 20 * 1e3:.3f} ms wall", flush=True)
dist.destroy_process_group()
