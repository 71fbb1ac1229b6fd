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


STAMP = os.path.join(PKG, "build", "sources.sha256")
FLAGS = ["-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC,-O3,-pthread"]


def _sha(paths, extra=""):
    import hashlib

    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(os.path.relpath(p, ROOT).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def source_hash():
    """sha256 over the contents of every source and header plus the compiler flags: the library is rebuilt whenever
    they differ from what the .so was built from (mtimes do not survive a snapshot copy to another box)."""
    return _sha(_deps(), " ".join(FLAGS))


def up_to_date():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_hash()


def build(force=False, verbose=False):
    """Compiles every .cu / .cpp under csrc/ for sm_100a and links libimp_b200.so.  ``force`` recompiles everything
    (what __graft_entry__.build() does); otherwise an object is reused only if the hash of its source, the headers and
    the flags matches the one recorded when it was compiled."""
    if up_to_date() and not force:
        return LIB
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    common = [NVCC, *FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    hdrs = _headers()
    for s in sources():
        o = os.path.join(PKG, "build", os.path.basename(s) + ".o")
        objs.append(o)
        want = _sha([s] + hdrs, " ".join(FLAGS))
        if not force and os.path.exists(o) and os.path.exists(o + ".sha256") and open(o + ".sha256").read().strip() == want:
            continue
        procs.append((s, o, want, subprocess.Popen(common + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, o, want, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        with open(o + ".sha256", "w") as f:
            f.write(want + "\n")
        if verbose or out.strip():
            print(out, file=sys.stderr)
    cmd = [NVCC, "-shared", *ARCH, "-o", LIB, *objs, "-Xcompiler", "-pthread", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
