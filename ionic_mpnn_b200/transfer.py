"""The transfer-learning recipe of the reference (train_melting_point_transfer.py) on the B200 kernels.

``build_transfer_model`` (:76-106): the trained viscosity model is cut at ``mix_cat_an`` (the sum of the two towers'
``Dense(mixing_size, relu)`` projections, train_viscosity.py:197-201) and gets a new head

    Dense(256, relu) "mp_dense_1" -> BatchNormalization "mp_bn_1" -> Dense(128, relu) "mp_dense_2" -> Dropout(0.3) "mp_dropout"
    -> Dense(64, relu) "mp_dense_3" -> Dense(1) "melting_point"

trained with ``Huber(delta=1.0)`` and ``Adam(lr)`` (no gradient clipping) in two stages (:189-241): stage 1 freezes every layer
whose name does not start with ``mp_`` / is not ``melting_point``; stage 2 additionally unfreezes the layers whose NAME contains
one of ``UNFREEZE_KEYS`` -- the bond-matrix and GatedUpdate layers of the last two message steps of each tower.  ``compile()``
is the reference's ``model.compile``: a fresh optimizer state.

Layer names are Keras' names: the reference's explicit ones (``{tower}_bmm_{i}``) and, for the auto-named GatedUpdate layers,
the ones a fresh process assigns in creation order -- ``gated_update``, ``gated_update_1`` ... ``gated_update_{2S-1}``, cation
tower first (train_viscosity.py:176-184); the same assumption the reference's own ``UNFREEZE_KEYS`` makes.

Numerics: everything runs in ``libimp_b200.so`` (base: the staged fp32 forward / backward kernels of train.py; head:
csrc/transfer_head.cu); torch holds device memory.  Inference (``predict``) uses whatever forward path the base model is
configured for (the fused tcgen05 kernel for precision="fp16") up to the molecule sums, then fp32 Dense layers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .train import TOWERS, _stream

UNFREEZE_KEYS = ["cat_bmm_2", "cat_bmm_3", "an_bmm_2", "an_bmm_3", "gated_update_2", "gated_update_3", "gated_update_6",
                 "gated_update_7", "mix_cat_an"]  # train_melting_point_transfer.py:206-212
BN_MOMENTUM, BN_EPS = 0.99, 1e-3  # keras.layers.BatchNormalization defaults


def head_shapes(mix):
    return {"mp_dense_1.kernel": (mix, 256), "mp_dense_1.bias": (256,), "mp_bn_1.gamma": (256,), "mp_bn_1.beta": (256,),
            "mp_bn_1.moving_mean": (256,), "mp_bn_1.moving_variance": (256,), "mp_dense_2.kernel": (256, 128),
            "mp_dense_2.bias": (128,), "mp_dense_3.kernel": (128, 64), "mp_dense_3.bias": (64,), "melting_point.kernel": (64, 1),
            "melting_point.bias": (1,)}


NON_TRAINABLE = ("mp_bn_1.moving_mean", "mp_bn_1.moving_variance")


def head_default_init(mix, seed=0):
    """Keras defaults: glorot_uniform kernels, zero biases, BatchNormalization gamma 1 / beta 0 / moving mean 0 / variance 1."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in head_shapes(mix).items():
        if k.endswith(".kernel"):
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            out[k] = rng.uniform(-lim, lim, size=shp).astype(np.float32)
        elif k.endswith((".gamma", ".moving_variance")):
            out[k] = np.ones(shp, np.float32)
        else:
            out[k] = np.zeros(shp, np.float32)
    return out


def keras_layer_of(var, num_steps):
    """Keras layer name that owns a variable of the base model (see the module docstring)."""
    grp = var.split(".")[0]
    for ti, t in enumerate(TOWERS):
        if grp.startswith(f"{t}_gu_"):
            k = ti * num_steps + int(grp.rsplit("_", 1)[1])
            return "gated_update" if k == 0 else f"gated_update_{k}"
    if grp == "atom_emb":
        return "embedding"
    if grp == "bond_emb":
        return "embedding_1"
    return grp  # {t}_bmm_{i} are the reference's own names; the fp / mix Dense layers keep their structural names here


class TransferModel:
    """``build_transfer_model(viscosity_model_path)`` + the two-stage training loop's pieces."""

    def __init__(self, base, seed=0, dropout=0.3):
        import torch

        if base.spec["kind"] != "viscosity":
            raise ValueError("the transfer recipe starts from the viscosity model (train_melting_point_transfer.py:78)")
        self.base, self.dropout, self.device = base, float(dropout), base.device
        self.mix = base.spec["mixing_size"]
        self.names = list(head_shapes(self.mix))
        shapes = head_shapes(self.mix)
        sizes = [int(np.prod(shapes[k])) for k in self.names]
        self.off = dict(zip(self.names, np.concatenate([[0], np.cumsum(sizes)[:-1]]).tolist()))
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=self.device)
        self.params = {k: self.flat[self.off[k]: self.off[k] + n].view(*shapes[k]) for k, n in zip(self.names, sizes)}
        self.set_head_weights(head_default_init(self.mix, seed))
        self.trainable = set()
        self.freeze_base()
        self._opt = None
        self._ws = {}
        self.seed, self.step_count = int(seed), 0

    # -- weights ------------------------------------------------------------------------------------------------
    def set_head_weights(self, w):
        import torch

        for k, v in w.items():
            if tuple(v.shape) != tuple(self.params[k].shape):
                raise ValueError(f"{k}: shape {tuple(v.shape)} != {tuple(self.params[k].shape)}")
            self.params[k].copy_(torch.from_numpy(np.ascontiguousarray(v, np.float32)))

    def get_head_weights(self):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.params.items()}

    def load_keras(self, path):
        """A saved transfer model (or a plain viscosity model: the head then keeps its current weights)."""
        from . import keras_io

        config, data = keras_io.read_keras(path)
        kind, params, extra = keras_io.params_from_keras(config, data)
        cur = self.base.get_weights()
        self.base.set_weights({**{k: cur[k] for k in cur if k.startswith("head")}, **{k: v for k, v in params.items() if not k.startswith("head")}})
        order = {"mp_dense_1": ("kernel", "bias"), "mp_dense_2": ("kernel", "bias"), "mp_dense_3": ("kernel", "bias"),
                 "melting_point": ("kernel", "bias"), "mp_bn_1": ("gamma", "beta", "moving_mean", "moving_variance")}
        for layer, arrs in extra.items():
            if layer in order:
                self.set_head_weights({f"{layer}.{n}": a for n, a in zip(order[layer], arrs)})
        return self

    def save_keras(self, path, key_style="class_counter"):
        """``model.save`` of the transfer model (train_melting_point_transfer.py:243-251): base up to mix_cat_an + the head."""
        from . import keras_io

        w = {k: v for k, v in self.base.get_weights().items() if not k.startswith("head")}
        w.update(self.get_head_weights())
        keras_io.export_keras(path, self.base.spec, w, key_style=key_style, transfer_head=True)

    # -- trainable flags (layer.trainable by Keras layer name) ------------------------------------------------------
    def layer_names(self):
        S = self.base.spec["num_steps"]
        names = {v: keras_layer_of(v, S) for v in self.base.var_names}
        names.update({v: v.split(".")[0] for v in self.names})
        return names

    def freeze_base(self):
        """Stage 1 (train_melting_point_transfer.py:191-193): only ``mp_*`` and ``melting_point`` stay trainable."""
        self.trainable = {v for v in self.names if v not in NON_TRAINABLE}
        return self

    def unfreeze(self, keys=UNFREEZE_KEYS):
        """Stage 2 (:214-217): ``layer.trainable = True`` for every layer whose name CONTAINS one of ``keys``."""
        for v, lname in self.layer_names().items():
            if v not in NON_TRAINABLE and any(k in lname for k in keys):
                self.trainable.add(v)
        return self

    def compile(self):
        """``model.compile(optimizer=Adam(...))``: a fresh optimizer (both stages of the reference re-compile)."""
        self._opt = None
        return self

    # -- forward -------------------------------------------------------------------------------------------------
    def _buf(self, name, numel):
        import torch

        t = self._ws.get(name)
        if t is None or t.numel() < numel:
            t = self._ws[name] = torch.empty(max(1, numel), dtype=torch.float32, device=self.device)
        return t

    def _p(self, k):
        return self.params[k].data_ptr()

    def _mixed_forward(self, pooled, P, keep):
        """``fp = Dense(fp_size, relu)`` and ``Dense(mixing_size, relu)`` per tower, then ``mix_cat_an`` (train_viscosity.py:
        189-201), as separate imp_dense calls so that the activations are there for the backward pass."""
        b, s, sm = self.base, self.base.spec, _stream()
        d, fp, mix = s["atom_dim"], s["fp_size"], s["mixing_size"]
        act = {}
        for ti, t in enumerate(TOWERS):
            x = pooled.data_ptr() + 4 * ti * P * d
            f, m = self._buf(f"{t}_fp", P * fp), self._buf(f"{t}_mixp", P * mix)
            _lib.call("imp_dense", x, P, d, fp, b._ptr(f"{t}_fp.kernel"), b._ptr(f"{t}_fp.bias"), 1, f.data_ptr(), sm)
            _lib.call("imp_dense", f.data_ptr(), P, fp, mix, b._ptr(f"{t}_mix.kernel"), b._ptr(f"{t}_mix.bias"), 1, m.data_ptr(), sm)
            act[t] = (x, f, m)
        mixed = self._buf("mixed", P * mix)
        _lib.call("imp_add", act["cat"][2].data_ptr(), act["an"][2].data_ptr(), P * mix, mixed.data_ptr(), sm)
        return mixed, act

    def _head_forward(self, mixed, P, training, seed=0):
        sm = _stream()
        a1, bn, a2, dr, a3 = (self._buf(n, P * c) for n, c in (("a1", 256), ("bn", 256), ("a2", 128), ("dr", 128), ("a3", 64)))
        out = self._buf("out", P)
        sv = self._buf("bn_stats", 512)
        _lib.call("imp_dense", mixed.data_ptr(), P, self.mix, 256, self._p("mp_dense_1.kernel"), self._p("mp_dense_1.bias"), 1, a1.data_ptr(), sm)
        _lib.call("imp_batchnorm", a1.data_ptr(), P, 256, self._p("mp_bn_1.gamma"), self._p("mp_bn_1.beta"), self._p("mp_bn_1.moving_mean"),
                  self._p("mp_bn_1.moving_variance"), C.c_float(BN_MOMENTUM), C.c_float(BN_EPS), 1 if training else 0, bn.data_ptr(),
                  sv.data_ptr() if training else None, sv.data_ptr() + 4 * 256 if training else None, sm)
        _lib.call("imp_dense", bn.data_ptr(), P, 256, 128, self._p("mp_dense_2.kernel"), self._p("mp_dense_2.bias"), 1, a2.data_ptr(), sm)
        if training and self.dropout > 0:
            _lib.call("imp_dropout", a2.data_ptr(), P * 128, C.c_float(self.dropout), C.c_uint64(seed), dr.data_ptr(), sm)
        else:
            dr = a2
        _lib.call("imp_dense", dr.data_ptr(), P, 128, 64, self._p("mp_dense_3.kernel"), self._p("mp_dense_3.bias"), 1, a3.data_ptr(), sm)
        _lib.call("imp_dense", a3.data_ptr(), P, 64, 1, self._p("melting_point.kernel"), self._p("melting_point.bias"), 0, out.data_ptr(), sm)
        return out, dict(a1=a1, bn=bn, a2=a2, dr=dr, a3=a3, sv=sv)

    def forward_packed(self, batch):
        """Inference on a device-resident packed batch -> device tensor [P] (standardised melting points, as the reference
        trains on ``scale(y)``, train_melting_point_transfer.py:180-186)."""
        if batch.dev is None:
            batch.to(self.device)
        pooled = self.base.pooled_sums(batch)
        mixed, _ = self._mixed_forward(pooled, batch.n_pairs, keep=False)
        out, _ = self._head_forward(mixed, batch.n_pairs, training=False)
        return out[: batch.n_pairs].clone()

    def predict(self, x):
        import torch

        batch = self.base.pack(x)
        out = self.forward_packed(batch)
        torch.cuda.current_stream().synchronize()
        self.base.check_status()
        return out.cpu().numpy().reshape(-1, 1)

    # -- training --------------------------------------------------------------------------------------------------
    def _optimizer(self):
        import torch

        if self._opt is None:
            nb, nh = self.base.flat.numel(), self.flat.numel()
            self._opt = {"mb": torch.zeros(nb, device=self.device), "vb": torch.zeros(nb, device=self.device),
                         "mh": torch.zeros(nh, device=self.device), "vh": torch.zeros(nh, device=self.device),
                         "gh": torch.zeros(nh, device=self.device), "step": 0,
                         "norms": torch.zeros(2 * (len(self.base.var_names) + len(self.names)) + 8, device=self.device),
                         "loss": torch.zeros(1, device=self.device)}
        return self._opt

    def loss_and_grads(self, batch, delta=1.0, training=True, dropout_seed=None):
        """Huber loss (mean over the batch) and the gradients of every TRAINABLE variable: head gradients in
        ``self._opt['gh']`` (head layout), base gradients in the base model's flat bucket.  Returns (loss sum, predictions)."""
        b, s = self.base, self.base.spec
        d, S, fp, mix, P = s["atom_dim"], s["num_steps"], s["fp_size"], s["mixing_size"], batch.n_pairs
        if batch.dev is None:
            batch.to(self.device)
        if batch.dev_y is None:
            raise ValueError("training needs batch.target (standardised melting points)")
        if not training:
            raise _lib.ImpError("loss_and_grads differentiates the training-mode graph (BatchNormalization batch statistics)")
        opt = self._optimizer()
        sm = _stream()
        base_trainable = [v for v in b.var_names if v in self.trainable]
        gnn_trainable = [v for v in base_trainable if "_bmm_" in v or "_gu_" in v or v in ("atom_emb", "bond_emb")]
        kept = b._forward_kept(batch)  # fp32 staged forward with the tape (also when the base is frozen: same arithmetic)
        mixed, act = self._mixed_forward(kept["pooled"], P, keep=True)
        seed = (self.seed * 1000003 + self.step_count) if dropout_seed is None else int(dropout_seed)
        out, hk = self._head_forward(mixed, P, training=training, seed=seed)
        loss = opt["loss"]
        gout = self._buf("g_out", P)
        _lib.call("imp_huber", out.data_ptr(), batch.dev_y.data_ptr(), P, C.c_float(delta), C.c_float(1.0 / P), loss.data_ptr(), gout.data_ptr(), sm)
        gh = opt["gh"]
        gh.zero_()
        gp = lambda k: gh.data_ptr() + 4 * self.off[k]  # noqa: E731
        g3, g2, gbn, g1, gm = (self._buf(n, P * c) for n, c in (("g3", 64), ("g2", 128), ("gbn", 256), ("g1", 256), ("gmix", mix)))
        _lib.call("imp_dense_bwd", hk["a3"].data_ptr(), None, gout.data_ptr(), P, 64, 1, self._p("melting_point.kernel"), 0, g3.data_ptr(),
                  gp("melting_point.kernel"), gp("melting_point.bias"), sm)
        _lib.call("imp_dense_bwd", hk["dr"].data_ptr(), hk["a3"].data_ptr(), g3.data_ptr(), P, 128, 64, self._p("mp_dense_3.kernel"), 1,
                  g2.data_ptr(), gp("mp_dense_3.kernel"), gp("mp_dense_3.bias"), sm)
        if training and self.dropout > 0:
            _lib.call("imp_dropout", g2.data_ptr(), P * 128, C.c_float(self.dropout), C.c_uint64(seed), g2.data_ptr(), sm)
        _lib.call("imp_dense_bwd", hk["bn"].data_ptr(), hk["a2"].data_ptr(), g2.data_ptr(), P, 256, 128, self._p("mp_dense_2.kernel"), 1,
                  gbn.data_ptr(), gp("mp_dense_2.kernel"), gp("mp_dense_2.bias"), sm)
        _lib.call("imp_batchnorm_bwd", hk["a1"].data_ptr(), gbn.data_ptr(), P, 256, self._p("mp_bn_1.gamma"), hk["sv"].data_ptr(),
                  hk["sv"].data_ptr() + 4 * 256, g1.data_ptr(), gp("mp_bn_1.gamma"), gp("mp_bn_1.beta"), sm)
        need_base = bool(base_trainable)
        _lib.call("imp_dense_bwd", mixed.data_ptr(), hk["a1"].data_ptr(), g1.data_ptr(), P, mix, 256, self._p("mp_dense_1.kernel"), 1,
                  gm.data_ptr() if need_base else None, gp("mp_dense_1.kernel"), gp("mp_dense_1.bias"), sm)
        if need_base:  # mix_cat_an is a sum: the same gradient flows into both towers' projections
            G = b._train_state()["g"]
            b._train_state()["grad"].zero_()
            for ti, t in enumerate(TOWERS):
                x, f, m = act[t]
                gf = self._buf(f"g_{t}_fp", P * fp)
                need_pool = bool(gnn_trainable)
                _lib.call("imp_dense_bwd", f.data_ptr(), m.data_ptr(), gm.data_ptr(), P, fp, mix, b._ptr(f"{t}_mix.kernel"), 1, gf.data_ptr(),
                          G[f"{t}_mix.kernel"].data_ptr(), G[f"{t}_mix.bias"].data_ptr(), sm)
                _lib.call("imp_dense_bwd", x, f.data_ptr(), gf.data_ptr(), P, d, fp, b._ptr(f"{t}_fp.kernel"), 1,
                          kept["dpooled"].data_ptr() + 4 * ti * P * d if need_pool else None, G[f"{t}_fp.kernel"].data_ptr(),
                          G[f"{t}_fp.bias"].data_ptr(), sm)
            if gnn_trainable:
                steps = [int(v.split(".")[0].rsplit("_", 1)[1]) for v in gnn_trainable if "_bmm_" in v or "_gu_" in v]
                first = 0 if any(v in ("atom_emb", "bond_emb") for v in gnn_trainable) else min(steps)
                b._backward_base(batch, kept, occurrence_norms=False, first_step=first)
        return loss, out[:P]

    def train_step(self, batch, lr=1e-3, delta=1.0, beta1=0.9, beta2=0.999, eps=1e-7, dropout_seed=None):
        """One ``model.fit`` batch of either stage: Huber loss, Adam (no clipping) on the trainable variables only.
        Returns the mean Huber loss of the batch (device scalar) computed with the weights before the update."""
        import torch

        loss, _ = self.loss_and_grads(batch, delta=delta, training=True, dropout_seed=dropout_seed)
        opt = self._optimizer()
        opt["step"] += 1
        self.step_count += 1
        sm = _stream()

        def adam(flat, grad, m, v, names, offsets, sizes):
            runs, cur = [], []
            for k in names:  # contiguous runs of trainable variables -> one imp_clip_adam call each
                if k in self.trainable:
                    cur.append(k)
                elif cur:
                    runs.append(cur)
                    cur = []
            if cur:
                runs.append(cur)
            for run in runs:
                base_off = offsets[run[0]]
                offs = torch.tensor([offsets[k] - base_off for k in run] + [offsets[run[-1]] + sizes[run[-1]] - base_off],
                                    dtype=torch.int64, device=self.device)
                l2 = torch.zeros(len(run), dtype=torch.float32, device=self.device)
                _lib.call("imp_clip_adam", flat.data_ptr() + 4 * base_off, grad.data_ptr() + 4 * base_off, m.data_ptr() + 4 * base_off,
                          v.data_ptr() + 4 * base_off, offs.data_ptr(), l2.data_ptr(), len(run), opt["norms"].data_ptr(), C.c_float(0.0),
                          C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps), opt["step"], sm)

        shapes = head_shapes(self.mix)
        adam(self.flat, opt["gh"], opt["mh"], opt["vh"], self.names, self.off, {k: int(np.prod(shapes[k])) for k in self.names})
        b = self.base
        if any(v in self.trainable for v in b.var_names):
            sizes = {k: int(np.prod(tuple(b.params[k].shape))) for k in b.var_names}
            adam(b.flat, b._train_state()["grad"], opt["mb"], opt["vb"], b.var_names, b.var_off, sizes)
            b._tables_valid = False
        return (loss[0] / batch.n_pairs).clone()
