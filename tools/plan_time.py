"""Times imp_fused_plan and imp_mpnn_forward_fused_planned separately (debug aid).  python tools/plan_time.py [pairs]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import _lib, graph  # noqa: E402

if os.environ.get("IMP_LIB"):
    _lib.LIB_PATH = os.environ["IMP_LIB"]
from ionic_mpnn_b200.viscosity import build_model  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
batch, _, _ = graph.synth_batch(pairs, seed=1003)
batch.to("cuda")
cb = batch.to_compact("cuda")
m = build_model(124, 72, precision="fp16", fused=True)
for _ in range(3):
    m.forward_packed(batch)
real = _lib.call
ev = []


def timed(n, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    real(n, *a)
    e1.record()
    ev.append((n, e0, e1))


import ionic_mpnn_b200.model as mm  # noqa: E402

mm._lib.call = timed
for b in (batch, cb, batch, cb):
    m.forward_packed(b)
torch.cuda.synchronize()
acc = {}
for n, e0, e1 in ev:
    acc.setdefault(n, []).append(e0.elapsed_time(e1))
for k, v in acc.items():
    print(f"{k:36s} " + " ".join(f"{x:.3f}" for x in v), "ms")
plan = m._ws["fused_plan"]
hdr = plan[:32].view(torch.int32).cpu().tolist()
print("tiles", hdr[:2], "cap", hdr[2:4], "status", hdr[4], "rows/tile", batch.n_atoms / max(1, hdr[0] + hdr[1]))
