"""Phase timeline of the wide GatedUpdate kernel (debug aid): clock64 at the phase boundaries of CTA 0's first tiles.
   python tools/wide_timeline.py [pairs] [extra_tc_flags]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import _lib, graph  # noqa: E402
from ionic_mpnn_b200.model import MPNNModel, make_spec  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
spec = make_spec("viscosity", atom_dim=256, num_steps=1)
batch, _, _ = graph.synth_batch(pairs, seed=1007, n_min=40, n_max=120)
batch.to("cuda")
m = MPNNModel(spec, precision="fp16")
m.extra_tc_flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m.forward_packed(batch)
buf = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
import ctypes  # noqa: E402

dbg = ctypes.CDLL(_lib.LIB_PATH).imp_debug_wide_timeline  # debug hook, deliberately not in include/imp_b200.h
dbg.restype, dbg.argtypes = None, [ctypes.c_void_p]
dbg(buf.data_ptr())
m.forward_packed(batch)
torch.cuda.synchronize()
dbg(None)
t = buf.cpu().numpy().reshape(16, 8)
names = ["wait phase A", "EA (r*h)", "wait phase B", "E2 pass 1", "E2 pass 2"]
print("tile  " + "  ".join(f"{n:>13s}" for n in names) + "   tile total")
for k in range(2, 14):
    d = [t[k][i + 1] - t[k][i] for i in range(5)]
    print(f"{k:4d}  " + "  ".join(f"{x:13d}" for x in d) + f"   {t[k + 1][0] - t[k][0]:10d}")
