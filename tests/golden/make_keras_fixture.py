#!/usr/bin/env python
"""Generate tests/golden/*.keras by serialising the REFERENCE'S OWN build_model graph (executed under oracle/tf_shim.py).

    python tests/golden/make_keras_fixture.py          (needs /root/reference; the fixtures are committed)

Keras itself is not available, so the archive is written by ionic_mpnn_b200.hdf5_min / zipfile -- but WHAT is written comes
from the reference: ``build_model`` of train_viscosity.py / train_melting_point.py is executed, and the recorded layer
objects are walked the way Keras' saver walks ``model.layers``:
  * config.json: one entry per layer of the functional graph in creation order -- class name, the layer's own name (the
    reference's explicit names and the auto-generated ones: dense_3, gated_update_5, ...), inbound_nodes from the recorded
    graph (the slice ``conn[:, :, 1]`` of train_viscosity.py:180 appears as its own op layer, as in Keras);
  * model.weights.h5: ``layers/<key>[/<attribute>]/vars/<n>`` with <key> = snake-cased class name + per-class counter in
    model.layers order (Keras saving_lib), <attribute> = dense_z / dense_r / dense_h / layernorm for the layers nested in
    GatedUpdate (models/layers.py:133-139), <n> = the order of add_weight calls.
The weights are the ones stored in the matching tests/golden/*.npz, injected by structure (make_golden.inject), so the
test can compare what MPNNModel.load_keras recovers against them.  A second archive keyed by LAYER NAME covers the other
naming convention the reader accepts.
"""
from __future__ import annotations

import json
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from ionic_mpnn_b200 import hdf5_min, keras_io  # noqa: E402
from oracle import ref_model, tf_shim  # noqa: E402


def serialise(model, key_style):
    shim_layer = tf_shim.Layer
    nested = set()
    for l in model.layers:
        for v in vars(l).values():
            if isinstance(v, shim_layer):
                nested.add(id(v))
    top = [l for l in model.layers if id(l) not in nested]
    layers, datasets = [], {}
    for node in model.inputs:
        layers.append({"class_name": "InputLayer", "name": node.name, "config": {"name": node.name}, "inbound_nodes": []})
    slice_names = {}

    def source_name(sym):
        if sym.name is not None:
            return sym.name
        # an op on a tensor (conn[:, :, 1]): its own layer, fed by the named ancestor
        if id(sym) not in slice_names:
            n = len(slice_names)
            nm = "tf.__operators__.getitem" + (f"_{n}" if n else "")
            slice_names[id(sym)] = nm
            layers.append({"class_name": "SlicingOpLambda", "name": nm, "config": {"name": nm},
                           "inbound_nodes": [[[source_name(sym.parents[0]), 0, 0, {}]]]})
        return slice_names[id(sym)]

    counters = {}
    for l in top:
        if l.output is None:
            continue
        inbound = [[source_name(p), 0, 0, {}] for p in l.output.parents]
        cls = type(l).__name__
        layers.append({"class_name": cls, "name": l.name, "config": {"name": l.name}, "inbound_nodes": [inbound]})
    # keys: per-class counter over model.layers order = the order of the config entries
    for entry in layers:
        base = keras_io.snake(entry["class_name"])
        n = counters.get(base, 0)
        counters[base] = n + 1
        entry["_key"] = entry["name"] if key_style == "layer_name" else (base if n == 0 else f"{base}_{n}")
    by_name = {l.name: l for l in top}
    for entry in layers:
        l = by_name.get(entry["name"])
        key = entry.pop("_key")
        if l is None:
            continue
        for n, w in enumerate(l.weights.values()):
            datasets[f"layers/{key}/vars/{n}"] = np.asarray(w, np.float32)
        for attr, sub in vars(l).items():
            if isinstance(sub, shim_layer):
                for n, w in enumerate(sub.weights.values()):
                    datasets[f"layers/{key}/{attr}/vars/{n}"] = np.asarray(w, np.float32)
    return {"class_name": "Functional", "config": {"name": "model", "layers": layers}}, datasets


def make(golden_name, out_name, key_style):
    z = np.load(os.path.join(HERE, golden_name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    spec, kind = meta["spec"], meta["kind"]
    params = {k[2:]: z[k] for k in z.files if k.startswith("w.")}
    assert params, f"{golden_name} does not store its weights"
    script = "train_viscosity.py" if kind == "viscosity" else "train_melting_point.py"
    tf_shim.configure(np.float64, seed=0)
    ns = mg.reference_namespace(script)
    kw = dict(atom_dim=spec["atom_dim"], fp_size=spec["fp_size"], mixing_size=spec["mixing_size"], num_steps=spec["num_steps"])
    if kind == "viscosity":
        kw["bond_dim"] = spec["bond_dim"]
    model = ns["build_model"](spec["atom_vocab_size"], spec["bond_vocab_size"], **kw)
    for r in meta["records"]:
        for ion in ("cation", "anion"):
            r[ion]["edge_indices"] = [tuple(e) for e in r[ion]["edge_indices"]]
    model.predict(mg.reference_inputs(ns, meta["records"], with_T=(kind == "viscosity")))  # builds the nested layers
    mg.inject(model, spec, {k: v.astype(np.float64) for k, v in params.items()}, spec["num_steps"])
    config, datasets = serialise(model, key_style)
    path = os.path.join(HERE, out_name)
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as zf:
        zf.writestr("metadata.json", json.dumps({"keras_version": "2.12.0", "writer": "tests/golden/make_keras_fixture.py"}))
        zf.writestr("config.json", json.dumps(config))
        zf.writestr("model.weights.h5", hdf5_min.write_datasets(datasets))
    print(f"{out_name}: {len(config['config']['layers'])} layers, {len(datasets)} variables, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    make("visc_small", "visc_small.keras", "class_counter")
    make("visc_small", "visc_small_by_name.keras", "layer_name")
    make("mp_small", "mp_small.keras", "class_counter")
