"""Wide tensor path (csrc/wide_tc.cu): per-stage CUDA-event times and tensor rates (debug aid).
   python tools/wide_probe.py [pairs] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import _lib, graph  # noqa: E402
from ionic_mpnn_b200.model import MPNNModel, make_spec  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
spec = make_spec("viscosity", atom_dim=256, num_steps=S)
batch, _, _ = graph.synth_batch(pairs, seed=1007, n_min=40, n_max=120)
batch.to("cuda")
m = MPNNModel(spec, precision="fp16")
for _ in range(2):
    out = m.forward_packed(batch)
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(out).all()), "atoms", batch.n_atoms, "unique entries", batch.n_unique, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m.forward_packed(batch)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 3
N, E, d = batch.n_atoms, batch.n_edges, 256
alg = S * (2 * E + 12 * N) * d * d
hw = S * (16 + 12) * N * d * d
print(f"forward {t:.3f} ms  {pairs / t * 1e3:.0f} pairs/s  algorithmic {alg / t / 1e9:.1f} TFLOP/s  executed {hw / t / 1e9:.1f} TFLOP/s", flush=True)
m.wide_per_stage_calls = True
real = _lib.call
ev = []


def timed(name, *a):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    real(name, *a)
    a1.record()
    ev.append((name, a0, a1))


import ionic_mpnn_b200.model as mm  # noqa: E402

mm._lib.call = timed
m.forward_packed(batch)
mm._lib.call = real
torch.cuda.synchronize()
acc = {}
for n, a0, a1 in ev:
    acc.setdefault(n, []).append(a0.elapsed_time(a1))
flop = {"imp_wide_message": 16 * N * d * d, "imp_wide_gates": 8 * N * d * d, "imp_wide_candidate": 4 * N * d * d}
for n, v in acc.items():
    ms = sum(v) / len(v)
    extra = f"  {flop[n] / ms / 1e9:.1f} TFLOP/s executed" if n in flop else ""
    print(f"{n:24s} {len(v):3d} x {ms:8.3f} ms{extra}")
