"""Phase timing of the fused forward kernel (debug aid, not part of the product).

Builds a second copy of the library with -DFZ_PROFILE into tools/_prof/, runs the fused kernel on a synthetic batch
and prints, per thread class, the share of clock64 cycles spent in each phase of the tile loop.

    python tools/fused_phase_profile.py [pairs]
"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_prof")
PHASES_H2X = ["between tiles / group load", "index loads + ballots", "bar (sort 1)", "offsets + rowof", "bar (sort 2)",
              "entry staging", "embedding + prefetch", "bar (tile ready)", "step: loop top / acc init", "pool", "bar (after pool)",
              "Z half 0 build", "Z half 1 build", "wait GEMM1a (before 2nd store)", "Z half store (tcgen05.st + wait)",
              "bar after Z half 0", "bar after Z half 1", "MMA issue (warp 0)", "wait GEMM1", "agg/h -> operands",
              "bar before GEMM2", "wait GEMM2", "gates z, r*h", "bar before GEMM3", "wait GEMM3", "candidate/blend/LN",
              "bar end of step", "(after steps)"] + ["-"] * 4
N_PROF = 32
PHASES = ["tile:idle/next", "sort+embed", "step:loop-top", "Z build", "bar after Z", "GEMM1 wait", "epi0 (agg,h->A)",
          "bar after epi0", "GEMM2 wait", "epi1 (z, r*h)", "bar after epi1", "GEMM3 wait", "epi2a (blend,sums)",
          "bar LN exchange", "epi2b (LN, h)", "bar end of step", "pool", "bar after pool"]


def main():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libimp_b200_prof.so")
    srcs = sorted(glob.glob(os.path.join(ROOT, "ionic_mpnn_b200", "csrc", "*.cu")) +
                  glob.glob(os.path.join(ROOT, "ionic_mpnn_b200", "csrc", "*.cpp")))
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DFZ_PROFILE", "-shared",
           "-Xcompiler", "-fPIC,-O3,-pthread", "-I", os.path.join(ROOT, "include"), "-I",
           os.path.join(ROOT, "ionic_mpnn_b200", "csrc"), "-o", lib, *srcs, "-cudart", "static"]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.check_call(cmd)
    if "--build-only" in sys.argv:
        return
    import torch

    from ionic_mpnn_b200 import _lib, graph
    _lib.LIB_PATH = lib
    from ionic_mpnn_b200.viscosity import build_model

    pairs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 262144
    batch, _, _ = graph.synth_batch(pairs, seed=1003)
    batch.to("cuda")
    m = build_model(124, 72, precision="fp16", fused=True, num_steps=int(os.environ.get("FZ_STEPS", "4")))
    if "--gen2" in sys.argv:
        m.extra_tc_flags = _lib.TC_TWO_THREADS_PER_ROW
    m._ws["status"] = torch.zeros(3 * N_PROF * 2, dtype=torch.int32, device="cuda")
    for _ in range(2):
        m.forward_packed(batch)
    torch.cuda.synchronize()
    m._ws["status"].zero_()
    m.forward_packed(batch)
    torch.cuda.synchronize()
    prof = m._ws["status"].view(torch.int64).cpu().numpy().reshape(3, N_PROF)
    labels = PHASES if "--gen2" in sys.argv else PHASES_H2X
    for cls, name in enumerate(["u=0 (warp 0, issues MMAs)", "u=96 (warp 3)", "u=224 (warp 7)"]):
        tot = prof[cls].sum()
        if tot == 0:
            continue
        print(f"--- {name}: total {tot / 1e6:.1f} Mcycles over all CTAs/contexts")
        for i, ph in enumerate(labels + ["-"] * (N_PROF - len(labels))):
            if prof[cls][i]:
                print(f"   {ph:28s} {100.0 * prof[cls][i] / tot:6.2f} %")


if __name__ == "__main__":
    main()
