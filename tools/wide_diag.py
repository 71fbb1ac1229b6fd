"""Wide tensor path vs the fp32 staged kernels: error of the pooled molecule sums per molecule (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import graph  # noqa: E402
from ionic_mpnn_b200.model import MPNNModel, make_spec  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 700
S = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16"
spec = make_spec("viscosity", atom_dim=256, num_steps=S)
batch, _, _ = graph.synth_batch(pairs, seed=11, n_min=40, n_max=120)
ref = MPNNModel(spec, seed=5, precision="fp32")
want, inter = ref.forward_packed(batch, keep=True)
hS = inter["h"][S].double()
mol_ptr = torch.from_numpy(batch.host["mol_ptr"].astype(np.int64)).cuda()
seg = torch.repeat_interleave(torch.arange(2 * pairs, device="cuda"), mol_ptr[1:] - mol_ptr[:-1])
pooled_ref = torch.zeros(2 * pairs, 256, dtype=torch.float64, device="cuda").index_add_(0, seg, hS)
m = MPNNModel(spec, seed=5, precision=prec)
m.set_weights(ref.get_weights())
got = m.forward_packed(batch)
pooled = m._ws["pooled"][: 2 * pairs * 256].view(2 * pairs, 256).double()
err = (pooled - pooled_ref).abs().amax(dim=1) / pooled_ref.abs().amax(dim=1)
print("pooled rel err per molecule: median %.3e  p99 %.3e  max %.3e (mol %d)" % (err.median(), err.quantile(0.99), err.max(), int(err.argmax())))
e = ((got - want).abs() / want.abs().clamp(min=1.0))
print("prediction rel err: median %.3e max %.3e; |want| range %.3f..%.3f" % (e.median(), e.max(), want.abs().min(), want.abs().max()))
# per-atom state error after the last step, read back from the tile-packed fp32 state
N = batch.n_atoms
ws = m._ws["wide_ws"]
rows = (N + 255) // 256 * 256
h32 = ws[: rows * 1024].view(torch.float32).view(rows // 128, 64, 128, 4).permute(0, 2, 1, 3).reshape(rows, 256)[:N].double()
ea = (h32 - hS).abs().amax(dim=1) / hS.abs().amax(dim=1)
deg = torch.from_numpy(np.diff(batch.host["row_ptr"])).cuda()
print("atom state rel err: median %.3e p99 %.3e max %.3e (row %d, unique in-degree %d)" % (ea.median(), ea.quantile(0.99), ea.max(), int(ea.argmax()), int(deg[ea.argmax()])))
for dg in range(0, int(deg.max()) + 1):
    sel = deg == dg
    if sel.any():
        print(f"  unique in-degree {dg}: {int(sel.sum())} rows, median err {ea[sel].median():.3e}, max {ea[sel].max():.3e}")
