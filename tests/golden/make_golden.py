#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE in this container.

Run from the repo root (needs /root/reference, which does not exist on the GPU box -- that is why
the vectors are committed):

    python tests/golden/make_golden.py

What runs, and from where (paths relative to /root/reference):
  * ``models/layers.py`` -- imported unmodified (BondMatrixMessage, Reduce, GatedUpdate, GlobalSumPool,
    SliceParamA/B/C, ScaleTemperature, ComputeLogEta, AddTwoTensors).
  * ``build_model``, ``pad_sequences_1d``, ``preprocess_edges_and_bonds`` -- lifted by AST from
    ``train_viscosity.py`` and ``train_melting_point.py`` (their module level needs sklearn /
    matplotlib / a data directory, so the modules are not imported whole).
  * the ``+1`` shifts and the ``build_inputs`` dict are the few inline lines of ``main()``
    (train_viscosity.py:255-262,291-314), restated below.
TensorFlow is replaced by ``oracle/tf_shim.py`` (numpy).  Weights are drawn by
``oracle.ref_model.init_params`` (Keras default initialisers) and injected by structure.
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import tf_shim  # noqa: E402
from oracle import ref_model  # noqa: E402
from ionic_mpnn_b200 import synth  # noqa: E402


def lift(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def reference_namespace(script):
    tf = tf_shim.install()
    sys.path.insert(0, REF)
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[m]
    import models.layers as L  # the reference module, verbatim

    ns = {"np": np, "tf": tf, "keras": tf.keras, "Input": tf.keras.layers.Input,
          "Embedding": tf.keras.layers.Embedding, "Dense": tf.keras.layers.Dense,
          "Add": tf.keras.layers.Add, "Model": tf.keras.models.Model, "l2": tf.keras.regularizers.l2}
    for n in ("BondMatrixMessage", "GatedUpdate", "GlobalSumPool", "Reduce", "AddTwoTensors",
              "SliceParamA", "SliceParamB", "SliceParamC", "ScaleTemperature", "ComputeLogEta"):
        ns[n] = getattr(L, n)
    lift(os.path.join(REF, script), {"build_model", "pad_sequences_1d", "preprocess_edges_and_bonds"}, ns)
    return ns


def reference_inputs(ns, records, with_T):
    # train_viscosity.py:255-262 (+1 on ids only) and :288-314 (maxima over the data set, dict keys)
    cat_atoms = [[a + 1 for a in d["cation"]["atom_ids"]] for d in records]
    cat_bonds = [[b + 1 for b in d["cation"]["bond_ids"]] for d in records]
    cat_edges = [d["cation"]["edge_indices"] for d in records]
    an_atoms = [[a + 1 for a in d["anion"]["atom_ids"]] for d in records]
    an_bonds = [[b + 1 for b in d["anion"]["bond_ids"]] for d in records]
    an_edges = [d["anion"]["edge_indices"] for d in records]
    max_atoms = max(max(map(len, cat_atoms)), max(map(len, an_atoms)))
    max_edges = max(max(map(len, cat_edges)), max(map(len, an_edges)))
    ce, cb = ns["preprocess_edges_and_bonds"](cat_edges, cat_bonds, max_edges)
    ae, ab = ns["preprocess_edges_and_bonds"](an_edges, an_bonds, max_edges)
    x = {"cat_atom": ns["pad_sequences_1d"](cat_atoms, max_atoms), "cat_bond": cb, "cat_connectivity": ce,
         "an_atom": ns["pad_sequences_1d"](an_atoms, max_atoms), "an_bond": ab, "an_connectivity": ae}
    if with_T:
        x["temperature"] = np.array([d["T"] for d in records], np.float32)[:, None]
    return x


def inject(model, spec, params, S):
    """Structure-based weight injection (Keras auto-names are a global counter)."""
    top = model.layers  # layers created while build_model ran, in creation order
    emb = [l for l in top if type(l).__name__ == "Embedding"]
    emb[0].weights["embeddings"] = params["atom_emb"]
    emb[1].weights["embeddings"] = params["bond_emb"]
    gus = [l for l in top if type(l).__name__ == "GatedUpdate"]
    dense = [l for l in top if type(l).__name__ == "Dense"]
    names = ["cat_fp", "an_fp", "cat_mix", "an_mix"] + (["head"] if spec["kind"] == "viscosity" else ["head1", "head2"])
    assert len(dense) == len(names) and len(gus) == 2 * S
    for l, n in zip(dense, names):
        l.weights["kernel"], l.weights["bias"] = params[f"{n}.kernel"], params[f"{n}.bias"]
    for ti, t in enumerate(ref_model.TOWERS):
        for i in range(S):
            model.get_layer(f"{t}_bmm_{i}").weights["bond_transform"] = params[f"{t}_bmm_{i}.bond_transform"]
            gu = gus[ti * S + i]
            for g in ("dense_z", "dense_r", "dense_h"):
                getattr(gu, g).weights["kernel"] = params[f"{t}_gu_{i}.{g}.kernel"]
                getattr(gu, g).weights["bias"] = params[f"{t}_gu_{i}.{g}.bias"]
            gu.layernorm.weights["gamma"] = params[f"{t}_gu_{i}.layernorm.gamma"]
            gu.layernorm.weights["beta"] = params[f"{t}_gu_{i}.layernorm.beta"]
    return gus


def run_case(name, kind, records, spec_kw, init_kw, float_dtype=np.float64):
    script = "train_viscosity.py" if kind == "viscosity" else "train_melting_point.py"
    tf_shim.configure(float_dtype, seed=0)
    ns = reference_namespace(script)
    spec = ref_model.make_spec(kind, **spec_kw)
    kw = dict(atom_dim=spec["atom_dim"], fp_size=spec["fp_size"], mixing_size=spec["mixing_size"],
              num_steps=spec["num_steps"])
    if kind == "viscosity":
        kw["bond_dim"] = spec["bond_dim"]
    model = ns["build_model"](spec["atom_vocab_size"], spec["bond_vocab_size"], **kw)
    x = reference_inputs(ns, records, with_T=(kind == "viscosity"))
    model.predict(x)  # builds the lazily-created sub-layers
    params = {k: v.astype(float_dtype) for k, v in ref_model.init_params(spec, **init_kw).items()}
    S = spec["num_steps"]
    gus = inject(model, spec, params, S)
    out = model.predict(x, batch_size=32)  # the reference's own call (train_viscosity.py:366)
    # intermediates, one feed of the whole set
    fetch, keys = [], []
    for ti, t in enumerate(ref_model.TOWERS):
        for i in range(S):
            fetch += [model.get_layer(f"{t}_bmm_{i}").output, model.get_layer(f"{t}_reduce_{i}").output,
                      gus[ti * S + i].output]
            keys += [f"{t}_msg_{i}", f"{t}_agg_{i}", f"{t}_h_{i + 1}"]
    pools = [l for l in model.layers if type(l).__name__ == "GlobalSumPool"]
    fetch += [pools[0].output, pools[1].output]
    keys += ["cat_pool", "an_pool"]
    vals = model.run({k: np.asarray(v) for k, v in x.items()}, fetch)
    blob = {f"x.{k}": v for k, v in x.items()}
    blob.update({f"i.{k}": np.asarray(v) for k, v in zip(keys, vals)})
    blob["out"] = np.asarray(out)
    store_w = sum(v.size for v in params.values()) < 40000
    if store_w:
        blob.update({f"w.{k}": v for k, v in params.items()})
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(np.ascontiguousarray(params[k], dtype=np.float64).tobytes())
    meta = {"name": name, "kind": kind, "spec": spec, "init": init_kw, "weights_stored": store_w,
            "weights_sha256_f64": h.hexdigest(), "float": np.dtype(float_dtype).name, "records": records}
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: out[:3]={np.asarray(out)[:3, 0]}  -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    small = synth.make_records(5, seed=0, n_min=3, n_max=12)
    std = synth.make_records(6, seed=0, n_min=10, n_max=40)
    many = synth.make_records(40, seed=3, n_min=3, n_max=9)  # > 32 pairs: exercises predict's batching
    mp_small = synth.make_records(5, seed=1, n_min=3, n_max=12, label="mp")
    # viscosity, reference defaults (d=32, K=8, S=4), Keras default init: weights re-derived from the seed
    run_case("visc_default_init", "viscosity", std, {}, dict(seed=1))
    # same graph, "trained-like" weights and bond_transform x10 so the message path dominates
    run_case("visc_trained_like", "viscosity", std, {}, dict(seed=2, trained_like=True, bond_scale=10.0))
    # small dims, weights stored in the file
    run_case("visc_small", "viscosity", small, dict(atom_dim=8, bond_dim=4, fp_size=8, mixing_size=6, num_steps=3),
             dict(seed=3, trained_like=True, bond_scale=8.0))
    run_case("visc_batched_predict", "viscosity", many,
             dict(atom_dim=8, bond_dim=4, fp_size=8, mixing_size=6, num_steps=2),
             dict(seed=4, trained_like=True, bond_scale=8.0))
    # melting point graph (bond_dim = atom_dim**2), small dims stored, default dims by seed
    run_case("mp_small", "melting_point", mp_small, dict(atom_dim=8, fp_size=8, mixing_size=6, num_steps=3),
             dict(seed=5, trained_like=True, bond_scale=30.0))
    run_case("mp_default_dims", "melting_point", mp_small, {}, dict(seed=6, trained_like=True, bond_scale=30.0))
    # float32 run of the reference wiring (what TF itself computes in), loose check only
    run_case("visc_default_init_f32", "viscosity", std, {}, dict(seed=1), float_dtype=np.float32)


if __name__ == "__main__":
    main()
