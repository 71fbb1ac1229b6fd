"""Small invocations of every new kernel path, meant to be run under compute-sanitizer (debug aid):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py   (where the pool allows it; also a plain smoke run)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import graph, synth  # noqa: E402
from ionic_mpnn_b200.model import MPNNModel, make_spec  # noqa: E402

# wide tensor path: 3 pairs (tiles straddling the tower boundary, padding items of the cluster form)
spec = make_spec("viscosity", atom_dim=256, num_steps=2)
b, _, _ = graph.synth_batch(3, seed=5, n_min=40, n_max=120)
for flags in (0, 256, 64):
    m = MPNNModel(spec, seed=1, precision="fp16")
    m.extra_tc_flags = flags
    y = m.predict(b)
    assert np.isfinite(y).all()
    print("wide flags", flags, "ok", flush=True)
# staged tensor paths at d = 32 (grouped tcgen05 message GEMM, folded Reduce, 16-bit rows, pipelined and one-chunk forms)
b, _, _ = graph.synth_batch(300, seed=6)
for kind in ("viscosity", "melting_point"):
    for prec in ("fp16", "bf16"):
        for flags in (0, 128):
            m = MPNNModel(make_spec(kind), seed=1, precision=prec, fused=False)
            m.extra_tc_flags = flags
            if kind == "melting_point":
                b.dev_T = None
            y = m.predict(b)
            assert np.isfinite(y).all()
            y2, inter = m.forward_packed(b, keep=True)
        print(kind, prec, "ok", flush=True)
# fp32 grouped messages + register-blocked GatedUpdate + training step
recs = synth.make_records(70, seed=3, label="log_eta")
tb = graph.pack_records(recs, 72, label="log_eta")
m = MPNNModel(make_spec("viscosity"), seed=2)
l0 = float(m.train_step(tb))
l1 = float(m.train_step(tb))
assert l1 < l0
print("train ok", l0, l1, flush=True)
