"""Builds ``libimp_b200.so`` (the C-ABI library of include/imp_b200.h) in-tree with nvcc for sm_100a.

    python -m ionic_mpnn_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libimp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in _deps())


def build(force=False, verbose=False):
    if up_to_date() and not force:
        return LIB
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    common = [NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC,-O3,-pthread",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for s in sources():
        o = os.path.join(PKG, "build", os.path.basename(s) + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in
                                                    [s] + glob.glob(os.path.join(CSRC, "*.cuh")) +
                                                    glob.glob(os.path.join(ROOT, "include", "*.h"))):
            continue
        procs.append((s, subprocess.Popen(common + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose or out.strip():
            print(out, file=sys.stderr)
    cmd = [NVCC, "-shared", *ARCH, "-o", LIB, *objs, "-Xcompiler", "-pthread", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
