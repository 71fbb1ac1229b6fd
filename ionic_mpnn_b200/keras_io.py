"""Reading (and writing) the reference's ``.keras`` model archives without Keras.

The reference saves trained models with ``model.save("...keras")`` and re-loads them with ``load_model``
(train_viscosity.py:353-354, train_melting_point.py:330-331, train_melting_point_transfer.py:78-93).  A ``.keras`` file is
a zip of ``config.json`` (the functional graph: layers with ``class_name``, ``name`` and ``inbound_nodes``),
``metadata.json`` and ``model.weights.h5`` (read by hdf5_min: no h5py here).

Variables are matched to ``model.param_shapes`` BY STRUCTURE, never by Keras' global auto-name counters (the reference
names only its ``{tower}_bmm_{i}`` / ``{tower}_reduce_{i}`` / ``mix_cat_an`` / head-slice layers; ``gated_update_3`` or
``dense_7`` depend on how many layers the process created before):

  Embedding layers in creation order          -> atom_emb, bond_emb                  (train_viscosity.py:163-164)
  ``{t}_bmm_{i}``                              -> {t}_bmm_{i}.bond_transform           (:176-178)
  the GatedUpdate fed by ``{t}_reduce_{i}``    -> {t}_gu_{i}.dense_z/r/h, layernorm    (:182-184, models/layers.py:128-140)
  Dense after GlobalSumPool of the last update -> {t}_fp ; the Dense it feeds -> {t}_mix   (:187-198)
  Dense fed by ``mix_cat_an`` / by ``Add``      -> head (viscosity) / head1 -> head2 (melting point)

Inside ``model.weights.h5`` every layer that owns variables has a group ``<container>/<key>[/<sublayer>]/vars/<n>``.
Two conventions for ``<key>`` are accepted: the layer's name, or Keras' saving_lib key -- the snake-cased class name with
a per-class counter in ``model.layers`` order (``dense``, ``dense_1``, ...).  Nested layers (the three Dense and the
LayerNormalization of a GatedUpdate) are keyed by their attribute names (models/layers.py:133-139).
"""
from __future__ import annotations

import io
import json
import re
import zipfile

import numpy as np

from . import hdf5_min

TOWERS = ("cat", "an")


def snake(name):
    s = re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub(r"([a-z])([A-Z])", r"\1_\2", s).lower()


# ------------------------------------------------------------------------------------------------- reading
def read_keras(path):
    """-> (config dict, {h5 dataset path: array})."""
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        if "config.json" not in names:
            raise ValueError(f"{path}: no config.json (not a Keras v3 archive)")
        config = json.loads(z.read("config.json"))
        wname = next((n for n in names if n.endswith(".weights.h5") or n.endswith("variables.h5")), None)
        if wname is None:
            raise ValueError(f"{path}: no model.weights.h5")
        data = hdf5_min.read_datasets(z.read(wname))
    return config, data


def _layers(config):
    cfg = config.get("config", config)
    out = []
    for l in cfg["layers"]:
        name = l.get("name") or l["config"]["name"]
        out.append({"class": l["class_name"], "name": name, "inputs": _inbound_names(l.get("inbound_nodes", [])),
                    "config": l.get("config", {})})
    return out


def _inbound_names(nodes):
    """Layer names feeding the FIRST call of a layer.  Keras 2 writes [[[name, node, tensor, kwargs], ...]]; Keras 3 writes
    {"args": [... {"class_name": "__keras_tensor__", "config": {"keras_history": [name, node, tensor]}} ...]}."""
    found = []

    def rec(x):
        if isinstance(x, dict):
            if "keras_history" in x:
                found.append(x["keras_history"][0])
            else:
                for v in x.values():
                    rec(v)
        elif isinstance(x, (list, tuple)):
            if len(x) >= 3 and isinstance(x[0], str) and isinstance(x[1], int) and isinstance(x[2], int):
                found.append(x[0])
                for v in x[3:]:  # kwargs may carry further tensors
                    rec(v)
            else:
                for v in x:
                    rec(v)

    if nodes:
        rec(nodes[0])
    return found


def _var_groups(data):
    """{(key, sublayer or ''): [arrays in vars order]} from dataset paths '<container>/<key>[/<sub>]/vars/<n>'."""
    groups = {}
    for path, arr in data.items():
        parts = path.split("/")
        if len(parts) < 4 or parts[-2] != "vars" or not parts[-1].isdigit():
            continue
        key, sub = parts[1], "/".join(parts[2:-2])
        groups.setdefault((key, sub), {})[int(parts[-1])] = arr
    return {k: [v[i] for i in sorted(v)] for k, v in groups.items()}


def params_from_keras(config, data):
    """-> (kind, {structural name: float32 array}) for a viscosity or melting-point model archive."""
    layers = _layers(config)
    groups = _var_groups(data)
    keys = {k for k, _ in groups}
    counters, key_of = {}, {}
    for l in layers:  # saving_lib key: snake-cased class name + per-class counter in model.layers order
        base = snake(l["class"])
        n = counters.get(base, 0)
        counters[base] = n + 1
        cand = base if n == 0 else f"{base}_{n}"
        key_of[l["name"]] = l["name"] if l["name"] in keys else cand
    by_name = {l["name"]: l for l in layers}
    consumers = {}
    for l in layers:
        for src in l["inputs"]:
            consumers.setdefault(src, []).append(l)

    def vars_of(name, sub=""):
        g = groups.get((key_of[name], sub))
        if g is None:
            raise ValueError(f"layer {name!r} (key {key_of[name]!r}{', ' + sub if sub else ''}) has no variables in the weights file")
        return g

    def consumer(name, cls):
        c = [l for l in consumers.get(name, []) if l["class"] == cls]
        if not c:
            raise ValueError(f"no {cls} layer consumes {name!r}")
        return c[0]

    out = {}
    emb = [l for l in layers if l["class"] == "Embedding"]
    if len(emb) != 2:
        raise ValueError(f"expected 2 Embedding layers (atoms, bonds), found {len(emb)}")
    out["atom_emb"], out["bond_emb"] = vars_of(emb[0]["name"])[0], vars_of(emb[1]["name"])[0]
    fps = {}
    for t in TOWERS:
        i, last = 0, None
        while f"{t}_bmm_{i}" in by_name:
            out[f"{t}_bmm_{i}.bond_transform"] = vars_of(f"{t}_bmm_{i}")[0]
            gu = consumer(f"{t}_reduce_{i}", "GatedUpdate")
            for g in ("dense_z", "dense_r", "dense_h"):
                out[f"{t}_gu_{i}.{g}.kernel"], out[f"{t}_gu_{i}.{g}.bias"] = vars_of(gu["name"], g)[:2]
            out[f"{t}_gu_{i}.layernorm.gamma"], out[f"{t}_gu_{i}.layernorm.beta"] = vars_of(gu["name"], "layernorm")[:2]
            last, i = gu, i + 1
        if last is None:
            raise ValueError(f"no {t}_bmm_0 layer: not a model of train_viscosity.py / train_melting_point.py")
        fp = consumer(consumer(last["name"], "GlobalSumPool")["name"], "Dense")
        out[f"{t}_fp.kernel"], out[f"{t}_fp.bias"] = vars_of(fp["name"])[:2]
        mix = consumer(fp["name"], "Dense")
        out[f"{t}_mix.kernel"], out[f"{t}_mix.bias"] = vars_of(mix["name"])[:2]
        fps[t] = mix
    merge = [l for l in consumers.get(fps["cat"]["name"], []) if l["class"] in ("AddTwoTensors", "Add")]
    if not merge:
        raise ValueError("no AddTwoTensors / Add layer joins the two towers")
    head = consumer(merge[0]["name"], "Dense")
    hv = vars_of(head["name"])
    extra = {}
    if hv[0].shape[-1] == 3 and merge[0]["class"] == "AddTwoTensors":
        kind = "viscosity"
        out["head.kernel"], out["head.bias"] = hv[:2]
    else:
        nxt = [l for l in consumers.get(head["name"], []) if l["class"] == "Dense"]
        if nxt and merge[0]["class"] == "Add":
            kind = "melting_point"
            out["head1.kernel"], out["head1.bias"] = hv[:2]
            out["head2.kernel"], out["head2.bias"] = vars_of(nxt[0]["name"])[:2]
        else:  # a transfer model (train_melting_point_transfer.py:95-104): the head layers carry their own names
            kind = "transfer"
            for l in layers:
                if l["name"].startswith("mp_") or l["name"] == "melting_point":
                    g = groups.get((key_of[l["name"]], ""))
                    if g:
                        extra[l["name"]] = g
    return kind, {k: np.asarray(v, np.float32) for k, v in out.items()}, extra


def spec_from_params(kind, params):
    d = params["atom_emb"].shape[1]
    S = sum(1 for k in params if k.startswith("cat_bmm_"))
    return dict(kind="viscosity" if kind == "transfer" else kind, atom_vocab_size=params["atom_emb"].shape[0],
                bond_vocab_size=params["bond_emb"].shape[0], atom_dim=d, bond_dim=params["bond_emb"].shape[1],
                fp_size=params["cat_fp.kernel"].shape[1], mixing_size=params["cat_mix.kernel"].shape[1], num_steps=S)


# ------------------------------------------------------------------------------------------------- writing
def reference_graph(spec, key_style="class_counter", transfer_head=False):
    """The functional graph of the reference's build_model as Keras 2.12 writes it to config.json: layers in creation
    order with Keras' auto names, and for every layer that owns variables the (h5 key, sublayer, structural names)."""
    kind, S = spec["kind"], spec["num_steps"]
    layers, owners = [], []
    auto, cls_count = {}, {}

    def add(cls, name=None, inputs=(), own=None):
        base = snake(cls)
        if name is None:
            n = auto.get(base, 0)
            auto[base] = n + 1
            name = base if n == 0 else f"{base}_{n}"
        layers.append({"class_name": cls, "name": name, "config": {"name": name},
                       "inbound_nodes": [[[src, 0, 0, {}] for src in inputs]] if inputs else []})
        c = cls_count.get(base, 0)
        cls_count[base] = c + 1
        key = name if key_style == "layer_name" else (base if c == 0 else f"{base}_{c}")
        if own:
            for sub, names in own:
                owners.append((key, sub, names))
        return name

    ins = ["cat_atom", "cat_bond", "cat_connectivity", "an_atom", "an_bond", "an_connectivity"]
    if kind == "viscosity":
        ins.append("temperature")
    for n in ins:
        add("InputLayer", n)
    # Keras creates a layer object where the script constructs it: the two embeddings first, then, tower by tower, the
    # layers of encode(); a layer appears in model.layers once, in creation order.
    emb_a = add("Embedding", own=[("", ["atom_emb"])])
    emb_b = add("Embedding", own=[("", ["bond_emb"])])
    fp_names = {}
    for t in TOWERS:
        h, be = emb_a, emb_b
        for i in range(S):
            m = add("BondMatrixMessage", f"{t}_bmm_{i}", [h, be, f"{t}_connectivity"], own=[("", [f"{t}_bmm_{i}.bond_transform"])])
            sl = add("SlicingOpLambda", None, [f"{t}_connectivity"])
            r = add("Reduce", f"{t}_reduce_{i}", [m, sl, h])
            p = f"{t}_gu_{i}"
            h = add("GatedUpdate", None, [h, r],
                    own=[(g, [f"{p}.{g}.kernel", f"{p}.{g}.bias"]) for g in ("dense_z", "dense_r", "dense_h")] +
                        [("layernorm", [f"{p}.layernorm.gamma", f"{p}.layernorm.beta"])])
        pool = add("GlobalSumPool", None, [h, f"{t}_atom"])
        fp_names[t] = add("Dense", None, [pool], own=[("", [f"{t}_fp.kernel", f"{t}_fp.bias"])])
    mix = {t: add("Dense", None, [fp_names[t]], own=[("", [f"{t}_mix.kernel", f"{t}_mix.bias"])]) for t in TOWERS}
    if kind == "viscosity" and transfer_head:  # build_transfer_model, train_melting_point_transfer.py:95-104
        x = add("AddTwoTensors", "mix_cat_an", [mix["cat"], mix["an"]])
        x = add("Dense", "mp_dense_1", [x], own=[("", ["mp_dense_1.kernel", "mp_dense_1.bias"])])
        x = add("BatchNormalization", "mp_bn_1", [x],
                own=[("", ["mp_bn_1.gamma", "mp_bn_1.beta", "mp_bn_1.moving_mean", "mp_bn_1.moving_variance"])])
        x = add("Dense", "mp_dense_2", [x], own=[("", ["mp_dense_2.kernel", "mp_dense_2.bias"])])
        x = add("Dropout", "mp_dropout", [x])
        x = add("Dense", "mp_dense_3", [x], own=[("", ["mp_dense_3.kernel", "mp_dense_3.bias"])])
        add("Dense", "melting_point", [x], own=[("", ["melting_point.kernel", "melting_point.bias"])])
    elif kind == "viscosity":
        mx = add("AddTwoTensors", "mix_cat_an", [mix["cat"], mix["an"]])
        hd = add("Dense", None, [mx], own=[("", ["head.kernel", "head.bias"])])
        a, b, c = (add(f"SliceParam{x}", f"param_{x}", [hd]) for x in "ABC")
        st = add("ScaleTemperature", "scale_T", ["temperature"])
        add("ComputeLogEta", "log_eta", [a, b, st, c])
    else:
        mx = add("Add", None, [mix["cat"], mix["an"]])
        h1 = add("Dense", None, [mx], own=[("", ["head1.kernel", "head1.bias"])])
        add("Dense", None, [h1], own=[("", ["head2.kernel", "head2.bias"])])
    return layers, owners


def export_keras(path, spec, params, key_style="class_counter", container="layers", transfer_head=False):
    """Writes ``params`` (structural names; with ``transfer_head`` also the mp_* / melting_point variables) as a ``.keras``
    archive with the reference's graph (see reference_graph)."""
    layers, owners = reference_graph(spec, key_style, transfer_head)
    if transfer_head:  # the inputs of the transfer model are the first six of the base (train_melting_point_transfer.py:93)
        layers = [l for l in layers if l["name"] not in ("temperature", "scale_T")]
    datasets = {}
    for key, sub, names in owners:
        for n, pname in enumerate(names):
            datasets["/".join(x for x in (container, key, sub, "vars", str(n)) if x)] = np.asarray(params[pname], np.float32)
    config = {"class_name": "Functional", "config": {"name": "model", "layers": layers}}
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("metadata.json", json.dumps({"keras_version": "2.12.0", "writer": "ionic_mpnn_b200.keras_io"}))
        z.writestr("config.json", json.dumps(config))
        z.writestr("model.weights.h5", hdf5_min.write_datasets(datasets))
