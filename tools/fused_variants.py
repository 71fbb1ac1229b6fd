"""A/B builds of the fused forward kernel (debug aid, not part of the product).

    python tools/fused_variants.py build  name=-DF4_SYNC=1 name2=-DF4_PEEL=0,-DF4_SYNC=1 ...   (here, no GPU needed)
    python tools/fused_variants.py run [pairs]                                                (on the GPU box)

`build` recompiles only csrc/fused_fwd*.cu with the extra flags and links them against the package's other objects
into tools/_prof/lib_<name>.so; `run` times every such library (and the in-tree one) with tools/fused_steps_sweep.py.
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_prof")


def build(specs):
    from ionic_mpnn_b200 import build as B

    B.build()
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, fl = spec.partition("=")
        flags = [f for f in fl.split(",") if f]
        objs = []
        for s in B.sources():
            base = os.path.basename(s)
            if base.startswith("fused_fwd"):
                o = os.path.join(OUT, f"{name}_{base}.o")
                procs.append(subprocess.Popen([B.NVCC, *B.FLAGS, *flags, "-I", os.path.join(ROOT, "include"), "-I", B.CSRC, "-c", s, "-o", o]))
            else:
                o = os.path.join(B.PKG, "build", base + ".o")
            objs.append(o)
        procs.append((name, objs))
    pend = [p for p in procs if not isinstance(p, tuple)]
    for p in pend:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    for name, objs in [p for p in procs if isinstance(p, tuple)]:
        lib = os.path.join(OUT, f"lib_{name}.so")
        subprocess.check_call([B.NVCC, "-shared", *B.ARCH, "-o", lib, *objs, "-Xcompiler", "-pthread", "-cudart", "static"])
        print("built", lib)


def run(pairs, flags_list):
    libs = [("in-tree", "")] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(OUT, "lib_*.so")))]
    for name, lib in libs:
        for fl in flags_list:
            env = dict(os.environ, FZ_FLAGS=str(fl))
            if lib:
                env["IMP_LIB"] = lib
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fused_steps_sweep.py"), str(pairs)], env=env,
                               capture_output=True, text=True)
            print(f"== {name} flags={fl}\n{r.stdout}{r.stderr[-400:] if r.returncode else ''}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
        flags_list = [int(x) for x in os.environ.get("FZ_FLAGS_LIST", "0").split(",")]
        run(pairs, flags_list)
