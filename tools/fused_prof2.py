"""Phase timing of the planned fused forward (generation 6; debug aid, not part of the product).

    python tools/fused_prof2.py build      (here: compiles csrc/fused_fwd*.cu with -DF6_PHASE_PROF into tools/_prof/lib_prof6.so)
    python tools/fused_prof2.py run [pairs] [steps] [gen 6|8]    (on the GPU box)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "_prof", "lib_prof6.so")
PH = ["pooling (+ next record issue)", "wait plan record", "slot + embedding", "bar (tile ready)", "Z build (accumulate)",
      "wait GEMM1a", "Z store", "bar operands (Z)", "MMA issue GEMM1 / gates", "wait gate GEMM", "r gate", "bar operands (r*h)",
      "MMA issue cand + z gate", "wait cand GEMM", "cand / LayerNorm / pack", "bar end of step"]

if sys.argv[1] == "build":
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "fused_variants.py"), "build", "prof6=-DF6_PHASE_PROF"])
else:
    import numpy as np
    import torch

    from ionic_mpnn_b200 import _lib, graph
    _lib.LIB_PATH = LIB
    from ionic_mpnn_b200.viscosity import build_model

    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    gen = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    batch, _, _ = graph.synth_batch(pairs, seed=1003)
    batch.to("cuda")
    m = build_model(124, 72, precision="fp16", fused=True, num_steps=steps)
    m.fused_gen = gen
    for _ in range(2):
        m.forward_packed(batch)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)()
    lib = _lib.load()
    getattr(lib, f"imp_debug_f{gen}_prof")(out)
    m.forward_packed(batch)
    torch.cuda.synchronize()
    getattr(lib, f"imp_debug_f{gen}_prof")(out)
    prof = np.array(list(out), dtype=np.float64).reshape(2, 16)
    for cls, name in enumerate(["thread 0 (warp 0: issues the MMAs)", "thread 96 (warp 3)"]):
        tot = prof[cls].sum()
        print(f"--- {name}: {tot / 1e6:.1f} Mcycles over all contexts")
        for i, ph in enumerate(PH):
            print(f"   {ph:32s} {100.0 * prof[cls][i] / tot:6.2f} %")
