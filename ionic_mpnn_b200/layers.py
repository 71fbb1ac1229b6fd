"""Drop-in layer classes with the reference's names, constructor arguments, list-of-inputs call
signatures and weight layouts (models/layers.py), running on the B200 kernels.

The reference builds its graph per tower (``encode`` in train_viscosity.py:166-190), so these layers are
called per tower too.  What changes is the tensor layout:

    reference (padded)                      here (packed)
    atom_state   (B, N, d)                  [n_atoms_of_tower, d] device tensor
    bond_state   (B, E, K) per-edge rows    ``BondState``: the (V_b, K) embedding table (never expanded)
    connectivity (B, E, 2)                  ``TowerView`` of a PackedGraphBatch
    atom_ids     (B, N)                     the same ``TowerView``

``MPNNModel.predict`` (model.py) does not go through these classes: it runs both towers in one launch
per stage.  These exist so that code written against models/layers.py keeps working layer by layer, and
they call exactly the same C-ABI entry points.

Head pieces (ComputeLogEta, ScaleTemperature, SliceParamA/B/C, AddTwoTensors) are [P,1]-sized
element-wise glue kept for API completeness on device tensors; the fused K6 kernel
(imp_pool_head_visc) is what the model path uses for them.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .graph import PackedGraphBatch


def _torch():
    import torch

    return torch


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


class TowerView:
    """One tower (0 = cation, 1 = anion) of a device-resident PackedGraphBatch."""

    def __init__(self, batch: PackedGraphBatch, tower: int):
        if batch.dev is None:
            raise _lib.ImpError("batch must be on the device: batch.to('cuda')")
        self.batch, self.tower = batch, tower
        P = batch.n_pairs
        self.a0 = 0 if tower == 0 else batch.n_cat_atoms
        self.a1 = batch.n_cat_atoms if tower == 0 else batch.n_atoms
        self.e0 = int(batch.host["row_ptr"][self.a0])
        self.e1 = int(batch.host["row_ptr"][self.a1])
        self.m0 = tower * P
        self.n_atoms = self.a1 - self.a0
        self.n_unique = self.e1 - self.e0
        self.n_mols = P

    def c_struct(self):
        """Graph struct restricted to this tower: row_ptr / atom_id / mol_ptr are offset, entry arrays stay
        absolute (row_ptr values are absolute entry indices, col_src values absolute atom indices)."""
        b = self.batch
        vb = b.bond_vocab
        s0 = int(b.host["bucket_ptr"][self.tower * vb])
        return _lib.Graph(self.n_mols, self.n_atoms, self.n_atoms, self.n_unique, 0, vb,
                          b.dev["mol_ptr"].data_ptr() + 4 * self.m0, b.dev["atom_id"].data_ptr() + 4 * self.a0,
                          b.dev["row_ptr"].data_ptr() + 4 * self.a0, b.dev["col_src"].data_ptr(), b.dev["edge_bm"].data_ptr(),
                          b.dev["bucket_ptr"].data_ptr() + 4 * self.tower * vb, b.dev["bucket_perm"].data_ptr() + 4 * s0)

    def virtual_base(self, t, width, first):
        """Pointer p such that p[(first + i) * width] is t[i]: lets absolute indices address a tower-local tensor."""
        return C.c_void_p(t.data_ptr() - 4 * first * width)


class BondState:
    def __init__(self, table):
        self.table = table  # (V_b, K) device tensor


class Layer:
    _counters = {}

    def __init__(self, name=None, **kwargs):
        base = "".join("_" + c.lower() if c.isupper() and i else c.lower() for i, c in enumerate(type(self).__name__))
        k = Layer._counters.get(base, 0)
        Layer._counters[base] = k + 1
        self.name = name or (base if k == 0 else f"{base}_{k}")
        self.built = False
        self.weights = {}

    def add_weight(self, shape, initializer="glorot_uniform", name=None, seed=None):
        torch = _torch()
        rng = np.random.default_rng(seed)
        if initializer == "glorot_uniform":
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            lim = math.sqrt(6.0 / ((shape[-2] + shape[-1]) * rf))
            w = rng.uniform(-lim, lim, size=shape)
        elif initializer == "uniform":
            w = rng.uniform(-0.05, 0.05, size=shape)
        elif initializer == "ones":
            w = np.ones(shape)
        else:
            w = np.zeros(shape)
        t = torch.from_numpy(w.astype(np.float32)).cuda()
        self.weights[name] = t
        return t

    def set_weights(self, **named):
        torch = _torch()
        for k, v in named.items():
            if k not in self.weights or tuple(self.weights[k].shape) != tuple(np.shape(v)):
                raise ValueError(f"{self.name}: bad weight {k}")
            self.weights[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda()

    def build(self, input_shape=None):
        pass

    def get_config(self):
        return {"name": self.name}

    def __call__(self, inputs, **kw):
        if not self.built:
            self.build(None)
            self.built = True
        return self.call(inputs, **kw)


class Embedding(Layer):
    """keras.layers.Embedding(vocab, dim, mask_zero=False) as used at train_viscosity.py:163-164.  Called on
    a TowerView it returns the atoms' rows (K1); ``as_bond_state()`` hands the table to BondMatrixMessage."""

    def __init__(self, input_dim, output_dim, mask_zero=False, **kw):
        super().__init__(**kw)
        assert not mask_zero
        self.input_dim, self.output_dim = input_dim, output_dim
        self.add_weight((input_dim, output_dim), "uniform", "embeddings")
        self.built = True

    def call(self, view: TowerView):
        torch = _torch()
        out = torch.empty(view.n_atoms, self.output_dim, dtype=torch.float32, device="cuda")
        _lib.call("imp_embed_atoms", self.weights["embeddings"].data_ptr(), self.input_dim,
                  view.batch.dev["atom_id"].data_ptr() + 4 * view.a0, view.n_atoms, self.output_dim, out.data_ptr(), _stream())
        return out

    def as_bond_state(self):
        return BondState(self.weights["embeddings"])


class BondMatrixMessage(Layer):
    """models/layers.py:86-125.  ``call([atom_state, bond_state, connectivity]) -> messages`` ([Eu_tower, d]
    in CSR order, multiplicity folded in); ``aggregate([...]) -> aggregated`` is the fused form of the draft
    models/bond_matrix_message.py:37-65."""

    def __init__(self, atom_dim, bond_dim, **kwargs):
        super().__init__(**kwargs)
        self.atom_dim, self.bond_dim = atom_dim, bond_dim

    def build(self, input_shape=None):
        self.bond_transform = self.add_weight((self.bond_dim, self.atom_dim, self.atom_dim), "glorot_uniform", "bond_transform")

    def _table(self, bond_state, interleaved):
        torch = _torch()
        vb = bond_state.table.shape[0]
        tab = torch.empty(vb * self.atom_dim * self.atom_dim, dtype=torch.float32, device="cuda")
        W = (C.c_void_p * 1)(self.weights["bond_transform"].data_ptr())
        T = (C.c_void_p * 1)(tab.data_ptr())
        _lib.call("imp_bond_table", bond_state.table.data_ptr(), vb, self.bond_dim, self.atom_dim, 1, W,
                  None if interleaved else T, T if interleaved else None, _stream())
        return tab

    def call(self, inputs):
        torch = _torch()
        atom_state, bond_state, view = inputs
        d = self.atom_dim
        tab = self._table(bond_state, False)
        msg = torch.empty(view.n_unique, d, dtype=torch.float32, device="cuda")
        g = view.c_struct()
        # messages are written at absolute entry positions: give the kernel a virtual base
        _lib.call("imp_edge_messages", C.byref(g), view.virtual_base(atom_state, d, view.a0), d, tab.data_ptr(),
                  tab.data_ptr(), view.virtual_base(msg, d, view.e0), _stream())
        return msg

    def aggregate(self, inputs):
        torch = _torch()
        atom_state, bond_state, view = inputs
        d = self.atom_dim
        if not self.built:
            self.build(None)
            self.built = True
        tab = self._table(bond_state, True)
        agg = torch.empty(view.n_atoms, d, dtype=torch.float32, device="cuda")
        g = view.c_struct()
        _lib.call("imp_message_agg", C.byref(g), view.virtual_base(atom_state, d, view.a0), d, tab.data_ptr(), tab.data_ptr(),
                  agg.data_ptr(), _stream())
        return agg

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"atom_dim": self.atom_dim, "bond_dim": self.bond_dim})
        return cfg


class Reduce(Layer):
    """models/layers.py:52-83: ``call([messages, tgt_idx, atom_ref]) -> aggregated``; ``tgt_idx`` is the
    TowerView (its CSR rows are the targets)."""

    def call(self, inputs):
        torch = _torch()
        messages, view, atom_ref = inputs
        d = atom_ref.shape[1]
        agg = torch.empty(view.n_atoms, d, dtype=torch.float32, device="cuda")
        g = view.c_struct()
        _lib.call("imp_segment_sum", C.byref(g), view.virtual_base(messages, d, view.e0), d, agg.data_ptr(), _stream())
        return agg


class GatedUpdate(Layer):
    """models/layers.py:128-156.  Weights: dense_z / dense_r / dense_h kernels (2d, d) + biases, LayerNorm
    gamma / beta; dropout_rate must be 0 at inference (the reference never sets it)."""

    def __init__(self, atom_dim, dropout_rate=0.0, **kwargs):
        super().__init__(**kwargs)
        self.atom_dim, self.dropout_rate = atom_dim, dropout_rate

    def build(self, input_shape=None):
        d = self.atom_dim
        for g in ("dense_z", "dense_r", "dense_h"):
            self.add_weight((2 * d, d), "glorot_uniform", f"{g}.kernel")
            self.add_weight((d,), "zeros", f"{g}.bias")
        self.add_weight((d,), "ones", "layernorm.gamma")
        self.add_weight((d,), "zeros", "layernorm.beta")

    def call(self, inputs, training=None):
        torch = _torch()
        if training and self.dropout_rate:
            raise NotImplementedError("dropout is not on the reference's path (rate 0)")
        atom_state, agg = inputs
        n, d = atom_state.shape
        w = _lib.GruWeights(*(self.weights[k].data_ptr() for k in
                              ("dense_z.kernel", "dense_z.bias", "dense_r.kernel", "dense_r.bias", "dense_h.kernel",
                               "dense_h.bias", "layernorm.gamma", "layernorm.beta")))
        out = torch.empty_like(atom_state)
        _lib.call("imp_gated_update", atom_state.data_ptr(), agg.data_ptr(), n, n, d, C.byref(w), C.byref(w),
                  C.c_float(1e-3), out.data_ptr(), _stream())
        return out


class GlobalSumPool(Layer):
    """models/layers.py:159-164: ``call([atom_features, atom_ids])``; ``atom_ids`` is the TowerView."""

    def call(self, inputs):
        torch = _torch()
        h, view = inputs
        d = h.shape[1]
        out = torch.empty(view.n_mols, d, dtype=torch.float32, device="cuda")
        b = view.batch
        # mol_ptr values are absolute atom indices: virtual bases for h and atom_id
        _lib.call("imp_global_sum_pool", b.dev["mol_ptr"].data_ptr() + 4 * view.m0, b.dev["atom_id"].data_ptr(), view.n_mols,
                  view.virtual_base(h, d, view.a0), d, out.data_ptr(), _stream())
        return out


class Dense(Layer):
    """keras.layers.Dense(units, activation) as the reference uses it (train_viscosity.py:189 ``Dense(fp_size, relu)``,
    :197-198 ``Dense(mixing_size, relu)``, :204 ``Dense(3)``; train_melting_point.py:173,191-198): weights ``kernel
    (in, units)`` / ``bias (units)``, ``y = act(x . kernel + bias)`` (imp_dense).  ``kernel_regularizer`` is accepted
    and recorded (it only matters for the training loss, which model.train_step computes)."""

    def __init__(self, units, activation=None, kernel_regularizer=None, input_dim=None, **kw):
        super().__init__(**kw)
        if activation not in (None, "linear", "relu"):
            raise ValueError("Dense: the reference only uses relu and linear activations")
        self.units, self.activation, self.kernel_regularizer = units, activation, kernel_regularizer
        self.in_dim = input_dim
        if input_dim is not None:  # weights can be set before the first call
            self.build(None)
            self.built = True

    def build(self, input_shape=None):
        if self.in_dim is None:
            raise ValueError("Dense.build needs the input width (call the layer on a tensor)")
        self.add_weight((self.in_dim, self.units), "glorot_uniform", "kernel")
        self.add_weight((self.units,), "zeros", "bias")

    def __call__(self, x, **kw):
        if not self.built:
            self.in_dim = int(x.shape[-1])
            self.build(None)
            self.built = True
        return self.call(x)

    def call(self, x):
        torch = _torch()
        x = x.contiguous()
        rows = x.numel() // self.in_dim
        y = torch.empty(*x.shape[:-1], self.units, dtype=torch.float32, device=x.device)
        _lib.call("imp_dense", x.data_ptr(), rows, self.in_dim, self.units, self.weights["kernel"].data_ptr(),
                  self.weights["bias"].data_ptr(), 1 if self.activation == "relu" else 0, y.data_ptr(), _stream())
        return y

    def get_config(self):
        cfg = super().get_config()
        cfg.update({"units": self.units, "activation": self.activation})
        return cfg


# ---- [P,1]-sized glue of the viscosity head (models/layers.py:10-49) -------------------------------
class ComputeLogEta(Layer):
    def call(self, inputs):
        A, B, T, Cc = inputs
        return A + B / (T + Cc + 1e-6)


class ScaleTemperature(Layer):
    def call(self, t):
        return t / 100.0


class SliceParamA(Layer):
    def call(self, x):
        return x[:, 0:1]


class SliceParamB(Layer):
    def call(self, x):
        torch = _torch()
        return torch.clamp(torch.nn.functional.softplus(x[:, 1:2]), 0.0, 20.0)


class SliceParamC(Layer):
    def call(self, x):
        torch = _torch()
        return torch.clamp(torch.nn.functional.softplus(x[:, 2:3]), 0.1, 50.0)


class AddTwoTensors(Layer):
    def call(self, inputs):
        a, b = inputs
        return a + b
