"""Scratch diagnostic: which LBO/SBO convention does the hardware follow?  (run on the GPU box)"""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from test_gpu_tensor import run_selftest, _bf16_round
for kind in (0, 1):
    for swap in (0, 1):
        for (N, K) in ((32, 32), (64, 64)):
            try:
                A, B, D = run_selftest(N, K, kind, swap)
                if kind == 0:
                    want = _bf16_round(A).astype(np.float64) @ _bf16_round(B).astype(np.float64).T
                else:
                    want = A.astype(np.float64) @ B.astype(np.float64).T
                print(f"kind={kind} swap={swap} N={N} K={K}: max|err|={np.abs(D - want).max():.4g}  max|want|={np.abs(want).max():.3g}", flush=True)
            except Exception as e:
                print("kind", kind, "swap", swap, N, K, "EXC", e, flush=True)
