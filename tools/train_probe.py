"""Where does a multi-GPU training step spend its device time?  (debug aid)"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import graph  # noqa: E402
from ionic_mpnn_b200.model import MPNNModel, make_spec  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = MPNNModel(make_spec("viscosity"), device=f"cuda:{local}")
b, _, _ = graph.synth_batch(65536, seed=2003 + local)
import numpy as np
b.target = np.random.default_rng(7).normal(2.0, 1.0, size=65536).astype("float32")
b.to(f"cuda:{local}")
for _ in range(3):
    m.train_step(b)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
st = m._train_state()
acc = [0.0] * 4
for it in range(5):
    ev[0].record()
    sse, _ = m.loss_and_grads(b, global_batch=65536 * world)
    ev[1].record()
    if world > 1:
        dist.all_reduce(st["grad"])
    ev[2].record()
    if world > 1:
        dist.all_reduce(sse)
    ev[3].record()
    torch.cuda.synchronize()
    for i in range(3):
        acc[i] += ev[i].elapsed_time(ev[i + 1]) / 5
if local == 0:
    print(f"world {world}: loss_and_grads {acc[0]:.2f} ms, all_reduce(grad) {acc[1]:.3f} ms, all_reduce(sse) {acc[2]:.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
