"""Fused forward time as a function of num_steps: separates the per-tile overhead (sort, embedding, pooling) from the
per-step cost (debug aid).   python tools/fused_steps_sweep.py [pairs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import _lib, graph  # noqa: E402

if os.environ.get("IMP_LIB"):
    _lib.LIB_PATH = os.environ["IMP_LIB"]  # A/B runs against another build of the library
from ionic_mpnn_b200.viscosity import build_model  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
batch, _, _ = graph.synth_batch(pairs, seed=1003)
batch.to("cuda")
cbatch = batch.to_compact("cuda")
for S in (1, 4):
    m = build_model(124, 72, num_steps=S, precision="fp16", fused=True)
    m.extra_tc_flags = int(os.environ.get("FZ_FLAGS", "0"))
    for _ in range(3):
        m.forward_packed(batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.forward_packed(batch)
    e1.record()
    torch.cuda.synchronize()
    t_full = e0.elapsed_time(e1) / 5
    for _ in range(3):
        m.forward_packed(cbatch)
    e0.record()
    for _ in range(5):
        m.forward_packed(cbatch)
    e1.record()
    torch.cuda.synchronize()
    print(f"steps {S}: int32 CSR {t_full:.3f} ms, compact feed {e0.elapsed_time(e1) / 5:.3f} ms per forward of {pairs} pairs", flush=True)
