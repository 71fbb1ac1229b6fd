"""fp32 staged path vs the fp64 oracle on BASELINE configs[0] (1000 pairs): per-stage kernel choices (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import synth  # noqa: E402
from ionic_mpnn_b200.viscosity import build_model  # noqa: E402
from oracle import ref_inputs, ref_model  # noqa: E402

recs = synth.make_records(1000, seed=0)
spec = ref_model.make_spec("viscosity")
params = ref_model.init_params(spec, seed=1)
x = ref_inputs.build_inputs(recs)
want = ref_model.predict(spec, params, x, batch_size=32)
f32 = ref_model.predict(spec, params, x, dtype=torch.float32, batch_size=32)


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


print("torch fp32 CPU port vs fp64:", rel(f32, want), " |want| min/max", float(np.abs(want).min()), float(np.abs(want).max()))
for simt in (False, True):
    m = build_model(124, 72)
    m.set_weights(params)
    m.simt_messages = simt
    got = m.predict(recs)
    e = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    print(f"simt_messages={simt}: max rel err {e.max():.3e}  median {np.median(e):.3e}  p99 {np.quantile(e, 0.99):.3e}")
