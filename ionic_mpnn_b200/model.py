"""Host-side model: the reference's ``build_model`` graphs executed as one packed two-tower batch.

Mirrors ``train_viscosity.py:139-231`` (``build_model`` with the same keyword defaults) and
``train_melting_point.py:137-215``.  Weight inventory, names-by-structure and shapes follow SURVEY 8b:
two shared embeddings; per tower x step one ``bond_transform (K,d,d)`` and the eight GatedUpdate
variables; per tower ``Dense(fp)`` and ``Dense(mix)``; the head.  Nothing is shared between towers or
steps (train_viscosity.py:176-189).

Everything numeric happens in ``libimp_b200.so`` through ctypes (``_lib``); torch is used for device
memory and streams only.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .graph import PackedGraphBatch, pack_padded, pack_records
from .train import TrainMixin

TOWERS = ("cat", "an")
DEFAULT_FUSED_GEN = "6"  # planned fused forward: kernel generation (MPNNModel.planned_gen)


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def param_shapes(spec):
    d, K, S = spec["atom_dim"], spec["bond_dim"], spec["num_steps"]
    fp, mix = spec["fp_size"], spec["mixing_size"]
    shapes = {"atom_emb": (spec["atom_vocab_size"], d), "bond_emb": (spec["bond_vocab_size"], K)}
    for t in TOWERS:
        for i in range(S):
            shapes[f"{t}_bmm_{i}.bond_transform"] = (K, d, d)
            for g in ("dense_z", "dense_r", "dense_h"):
                shapes[f"{t}_gu_{i}.{g}.kernel"] = (2 * d, d)
                shapes[f"{t}_gu_{i}.{g}.bias"] = (d,)
            shapes[f"{t}_gu_{i}.layernorm.gamma"] = (d,)
            shapes[f"{t}_gu_{i}.layernorm.beta"] = (d,)
        shapes[f"{t}_fp.kernel"] = (d, fp)
        shapes[f"{t}_fp.bias"] = (fp,)
    for t in TOWERS:
        shapes[f"{t}_mix.kernel"] = (fp, mix)
        shapes[f"{t}_mix.bias"] = (mix,)
    if spec["kind"] == "viscosity":
        shapes["head.kernel"] = (mix, 3)
        shapes["head.bias"] = (3,)
    else:
        shapes["head1.kernel"] = (mix, fp)
        shapes["head1.bias"] = (fp,)
        shapes["head2.kernel"] = (fp, 1)
        shapes["head2.bias"] = (1,)
    return shapes


def keras_default_init(spec, seed=0):
    """Keras default initialisers: Embedding U(-0.05,0.05); glorot_uniform kernels (rank-3
    ``bond_transform``: fans multiplied by K); zero biases; LayerNorm gamma 1, beta 0."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in param_shapes(spec).items():
        leaf = name.split(".")[-1]
        if name in ("atom_emb", "bond_emb"):
            w = rng.uniform(-0.05, 0.05, size=shp)
        elif leaf in ("kernel", "bond_transform"):
            rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            lim = math.sqrt(6.0 / ((shp[-2] + shp[-1]) * rf))
            w = rng.uniform(-lim, lim, size=shp)
        elif leaf == "gamma":
            w = np.ones(shp)
        else:
            w = np.zeros(shp)
        out[name] = w.astype(np.float32)
    return out


class MPNNModel(TrainMixin):
    """Object returned by ``build_model``: ``predict(x)`` like the Keras model, on the B200 kernels."""

    LN_EPS = 1e-3  # Keras LayerNormalization default (models/layers.py:139)

    PRECISIONS = ("fp32", "bf16", "bf16_precise", "fp16", "fp16_precise")

    def __init__(self, spec, device="cuda", seed=0, precision="fp32", fused="auto"):
        """``precision``: 'fp32' = SIMT kernels (the 1e-5 path); 'fp16' / 'bf16' = tcgen05 tensor-core path with
        IEEE-half / bfloat16 operands and fp32 accumulation (the 2e-2 path; '*_precise' keeps expf/tanhf in the
        epilogues).  ``fused``: run the whole-tower fused kernel (imp_mpnn_forward_fused) when the shape allows it
        ('auto' = yes for the tensor precisions), else the staged per-layer kernels.

        Tuning attributes (plain attributes, all optional): ``fused_gen`` (6 default; 7 / 8 = the measured alternatives of the
        planned fused forward), ``plan_slack`` ('auto', a factor, or None = never-overflowing plan buffer),
        ``use_plan`` (False = the self-contained fused kernel), ``fp32_tensor`` (precision 'fp32' only: 3xTF32 tensor-core
        GatedUpdate and messages -- 1.57x, fp32-class but 1.6e-5 on the worst element, NOT the 1e-5 path), ``tc_forward`` /
        ``tc_backward`` (training: False = the fp32 SIMT kernels), ``extra_tc_flags`` (IMP_TC_* experiment switches)."""
        import torch

        if precision not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {self.PRECISIONS}")
        self.precision = precision
        self.fused = fused

        _lib.load()
        if not torch.cuda.is_available():
            raise _lib.ImpError("ionic_mpnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.spec = dict(spec)
        self.device = torch.device(device)
        self.params = {}
        self.flat = None
        self.set_weights(keras_default_init(spec, seed))
        self._ws = {}
        self._tables_valid = False

    # -- weights ----------------------------------------------------------------------------
    def set_weights(self, weights):
        import torch

        shapes = param_shapes(self.spec)
        if getattr(self, "flat", None) is None:
            # one flat fp32 buffer (the gradient bucket mirrors it); every variable is a view, 16-byte aligned
            self.var_names, self.var_off, off = list(shapes), {}, 0
            for k, shp in shapes.items():
                self.var_off[k] = off
                off += (int(np.prod(shp)) + 3) // 4 * 4
            self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
            for k, shp in shapes.items():
                self.params[k] = self.flat[self.var_off[k]: self.var_off[k] + int(np.prod(shp))].view(*shp)
        for k, shp in shapes.items():
            if k not in weights:
                raise KeyError(f"missing weight {k}")
            w = np.ascontiguousarray(np.asarray(weights[k], dtype=np.float32))
            if tuple(w.shape) != tuple(shp):
                raise ValueError(f"{k}: shape {w.shape} != {shp}")
            self.params[k].copy_(torch.from_numpy(w))
        self._tables_valid = False

    def get_weights(self):
        return {k: v.detach().cpu().numpy() for k, v in self.params.items()}

    def save_weights(self, path, include_optimizer=True):
        """Checkpoint (``model.save`` role, train_viscosity.py:353-354): one ``.npz`` keyed by the variable names of
        ``param_shapes`` (INTEGRATION.md maps them to the reference's Keras layers) plus the model spec and, when a
        training state exists, Adam's m / v / step so that ``fit`` / ``train_step`` resume bit-identically."""
        import json

        arrays = {f"w/{k}": v for k, v in self.get_weights().items()}
        arrays["spec"] = np.frombuffer(json.dumps(self.spec, sort_keys=True).encode(), dtype=np.uint8)
        st = getattr(self, "_train", None)
        if include_optimizer and st is not None:
            arrays["adam/m"], arrays["adam/v"] = st["m"].cpu().numpy(), st["v"].cpu().numpy()
            arrays["adam/step"] = np.asarray(st["step"], dtype=np.int64)
        with open(path, "wb") as f:
            np.savez(f, **arrays)

    def load_weights(self, path):
        """Inverse of save_weights.  The file's spec must match this model's (shapes are checked variable by variable)."""
        import json

        import torch

        with np.load(path) as z:
            if "spec" in z.files:
                spec = json.loads(bytes(z["spec"]).decode())
                diff = {k: (spec.get(k), self.spec.get(k)) for k in self.spec if spec.get(k) != self.spec.get(k)}
                if diff:
                    raise ValueError(f"checkpoint was written for a different model spec: {diff}")
            self.set_weights({k[2:]: z[k] for k in z.files if k.startswith("w/")})
            if "adam/m" in z.files:
                st = self._train_state()
                st["m"].copy_(torch.from_numpy(z["adam/m"]))
                st["v"].copy_(torch.from_numpy(z["adam/v"]))
                st["step"] = int(z["adam/step"])
        return self

    def load_keras(self, path):
        """Loads the variables of a reference ``.keras`` archive (``model.save``, train_viscosity.py:353-354) into this model.
        Variables are matched by graph structure (keras_io.params_from_keras); every shape is checked against the spec."""
        from . import keras_io

        config, data = keras_io.read_keras(path)
        kind, params, _ = keras_io.params_from_keras(config, data)
        want = "viscosity" if kind == "transfer" else kind
        if want != self.spec["kind"]:
            raise ValueError(f"{path} holds a {kind} model, this is a {self.spec['kind']} model")
        if kind == "transfer":  # the base of a transfer model has no Dense(3) head: keep ours
            params = {k: v for k, v in params.items() if not k.startswith("head")}
            cur = self.get_weights()
            params = {**{k: cur[k] for k in cur if k.startswith("head")}, **params}
        self.set_weights(params)
        return self

    def save_keras(self, path, key_style="class_counter"):
        """Writes the variables as a ``.keras`` archive with the reference's functional graph in config.json (zip of
        config.json + model.weights.h5; HDF5 written by hdf5_min, see there for what has and has not been verified)."""
        from . import keras_io

        keras_io.export_keras(path, self.spec, self.get_weights(), key_style=key_style)

    def count_params(self):
        return sum(int(np.prod(s)) for s in param_shapes(self.spec).values())

    # -- workspace ---------------------------------------------------------------------------
    def _buf(self, name, numel, dtype=None):
        import torch

        dtype = dtype or torch.float32
        t = self._ws.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(max(numel, 1), dtype=dtype, device=self.device)
            self._ws[name] = t
        return t

    def _ptr(self, name):
        return C.c_void_p(self.params[name].data_ptr())

    def _gru_struct(self, t, i):
        p = f"{t}_gu_{i}"
        return _lib.GruWeights(*(self.params[f"{p}.{n}"].data_ptr() for n in
                                 ("dense_z.kernel", "dense_z.bias", "dense_r.kernel", "dense_r.bias", "dense_h.kernel",
                                  "dense_h.bias", "layernorm.gamma", "layernorm.beta")))

    def _readout_struct(self, t):
        return _lib.ReadoutWeights(*(self.params[n].data_ptr() for n in
                                     (f"{t}_fp.kernel", f"{t}_fp.bias", f"{t}_mix.kernel", f"{t}_mix.bias")))

    def refresh_tables(self):
        """K2 for all (tower, step) in one launch; called once per weight set (once per step in training)."""
        s = self.spec
        S, d, Vb, K = s["num_steps"], s["atom_dim"], s["bond_vocab_size"], s["bond_dim"]
        n = 2 * S
        if self.wide_supported():
            # wide tensor path (csrc/wide_tc.cu): bond tables are never formed (messages are Z . Wc), one packed block
            # per (tower, step)
            import torch

            wb = _lib.load().imp_wide_pack_bytes(d, K)
            wpk = self._buf("wide_packed", wb * n, torch.uint8)
            for ti, t in enumerate(TOWERS):
                for i in range(S):
                    w = self._gru_struct(t, i)
                    _lib.call("imp_wide_pack", self._ptr(f"{t}_bmm_{i}.bond_transform"), C.byref(w), d, K,
                              wpk.data_ptr() + wb * (ti * S + i), _stream())
            self._tables_valid = True
            return
        per = Vb * d * d
        tab = self._buf("table", n * per)
        tab_il = self._buf("table_il", n * per)
        W = (C.c_void_p * n)(*[self.params[f"{t}_bmm_{i}.bond_transform"].data_ptr() for t in TOWERS for i in range(S)])
        T0 = (C.c_void_p * n)(*[tab.data_ptr() + 4 * per * j for j in range(n)])
        T1 = (C.c_void_p * n)(*[tab_il.data_ptr() + 4 * per * j for j in range(n)])
        _lib.call("imp_bond_table", self._ptr("bond_emb"), Vb, K, d, n, W, T0, T1, _stream())
        if self.precision != "fp32":
            import torch

            nbytes = _lib.load().imp_gru_pack_bytes(d)
            if nbytes < 0:
                raise _lib.ImpError(f"tensor path does not support atom_dim {d}")
            # tensor-core message kernel (csrc/msg_tc.cu): the tables as UMMA operand slices
            mb = _lib.load().imp_message_pack_bytes(Vb, d)
            self._msg_pack_bytes = (mb + 255) // 256 * 256
            mpk = self._buf("msg_packed", self._msg_pack_bytes * n, torch.uint8)
            for j in range(n):
                _lib.call("imp_message_pack", tab.data_ptr() + 4 * per * j, Vb, d, self.tc_flags(),
                          mpk.data_ptr() + self._msg_pack_bytes * j, _stream())
            self._gru_pack_bytes = (nbytes + 255) // 256 * 256
            pk = self._buf("gru_packed", self._gru_pack_bytes * n, torch.uint8)
            pack_fn = "imp_gru_pack_f16" if self.precision.startswith("fp16") else "imp_gru_pack_bf16"
            for ti, t in enumerate(TOWERS):
                for i in range(S):
                    w = self._gru_struct(t, i)
                    _lib.call(pack_fn, C.byref(w), d, pk.data_ptr() + self._gru_pack_bytes * (ti * S + i), _stream())
            if self.fused_supported():
                fb = _lib.load().imp_fused_pack_bytes(d, K)
                fpk = self._buf("fused_packed", fb * n, torch.uint8)
                for ti, t in enumerate(TOWERS):
                    for i in range(S):
                        w = self._gru_struct(t, i)
                        _lib.call("imp_fused_pack", self._ptr(f"{t}_bmm_{i}.bond_transform"), C.byref(w), d, K,
                                  self.tc_flags() & ~_lib.TC_GEN5, fpk.data_ptr() + fb * (ti * S + i), _stream())
                if self.planned_supported():  # the planned forward's own operand layout (tf32 blocks for the agg terms)
                    fb6 = _lib.load().imp_fused_pack_planned_bytes(d, K)
                    fpk6 = self._buf("fused_packed6", fb6 * n, torch.uint8)
                    for ti, t in enumerate(TOWERS):
                        for i in range(S):
                            w = self._gru_struct(t, i)
                            _lib.call("imp_fused_pack_planned7" if self.planned_gen() in (7, 8) else "imp_fused_pack_planned",
                                      self._ptr(f"{t}_bmm_{i}.bond_transform"), C.byref(w), d, K,
                                      fpk6.data_ptr() + fb6 * (ti * S + i), _stream())
                    self._packed6_gen = self.planned_gen()
        self._tables_valid = True

    def planned_gen(self):
        """Kernel generation of the planned fused forward: 7 (csrc/fused_fwd7.cu) or 6 (csrc/fused_fwd6.cu); 5 through
        ``extra_tc_flags = TC_GEN5``.  ``fused_gen`` on the model or IMP_FUSED_GEN in the environment select it."""
        import os

        if self.tc_flags() & _lib.TC_GEN5:
            return 5
        return int(getattr(self, "fused_gen", None) or os.environ.get("IMP_FUSED_GEN", DEFAULT_FUSED_GEN))

    def tc_flags(self):
        f = _lib.TC_FP16 if self.precision.startswith("fp16") else 0
        if self.precision.endswith("_precise"):
            f |= _lib.TC_PRECISE_EPILOGUE
        return f | getattr(self, "extra_tc_flags", 0)

    def wide_supported(self):
        """Shape envelope of imp_mpnn_forward_wide (include/imp_b200.h): IEEE-half operands, atom_dim 256, bond_dim 8."""
        s = self.spec
        return (self.precision.startswith("fp16") and s["atom_dim"] == 256 and s["bond_dim"] == 8
                and s["bond_vocab_size"] <= 256)

    def fused_supported(self):
        """Shape envelope of imp_mpnn_forward_fused (include/imp_b200.h)."""
        s = self.spec
        return (self.precision != "fp32" and s["atom_dim"] == 32 and s["bond_dim"] == 8 and 1 <= s["num_steps"] <= 4
                and s["bond_vocab_size"] <= 256)

    def compact_supported(self):
        """The compact input feed is read by the default half-operand fused kernel (no tuning flags), 8-bit vocabularies."""
        f = self.tc_flags()
        return (self.fused_supported() and (f & _lib.TC_FP16) and not (f & (_lib.TC_F32_ZBUILD | _lib.TC_TWO_THREADS_PER_ROW))
                and self.spec["atom_vocab_size"] <= 256 and self.spec["bond_vocab_size"] <= 256)

    def planned_supported(self, batch=None):
        """The planned (fifth-generation) fused forward: IEEE-half operands, no kernel-generation tuning flags, and a batch
        inside the tile plan's envelope (in-degree <= 31, <= 336 unique entries per molecule; a device-packed batch does
        not know these on the host -- the plan's status word, read by check_status(), reports a violation)."""
        f = self.tc_flags()
        if not (self.fused_supported() and (f & _lib.TC_FP16) and not (f & ~(_lib.TC_FP16 | _lib.TC_PRECISE_EPILOGUE | _lib.TC_GEN5))
                and self.spec["atom_vocab_size"] <= 1024 and getattr(self, "use_plan", True)):
            return False
        if batch is not None:
            deg, ents = getattr(batch, "max_in_degree", None), getattr(batch, "max_mol_entries", None)
            if (deg is not None and deg > 31) or (ents is not None and ents > 336):
                return False
        return True

    def use_fused(self, batch):
        if self.fused is False or not self.fused_supported():
            if self.fused is True:
                raise _lib.ImpError("fused=True, but the model shape is outside the fused kernel's envelope "
                                    "(tensor precision, atom_dim 32, bond_dim 8, <= 4 steps)")
            return False
        if batch.max_mol_atoms > 128:
            if self.fused is True:
                raise _lib.ImpError(f"fused=True, but a molecule has {batch.max_mol_atoms} atoms (tile = 128)")
            return False
        return True

    def table_ptr(self, tower, step, interleaved):
        s = self.spec
        per = s["bond_vocab_size"] * s["atom_dim"] ** 2
        j = tower * s["num_steps"] + step
        base = self._ws["table_il" if interleaved else "table"].data_ptr()
        return C.c_void_p(base + 4 * per * j)

    # -- forward -----------------------------------------------------------------------------
    def forward_packed(self, batch: PackedGraphBatch, keep=False, unfused_messages=False):
        """Runs the whole graph on a device-resident packed batch.  Returns a device tensor [P] (and, with
        ``keep``, a dict of device tensors of every intermediate).  ``unfused_messages`` routes the message
        step through K3 (imp_edge_messages) + K4 (imp_segment_sum) instead of the fused imp_message_agg."""
        import torch

        s = self.spec
        d, S = s["atom_dim"], s["num_steps"]
        if batch.dev is None:
            batch.to(self.device)
        if batch.bond_vocab != s["bond_vocab_size"]:
            raise ValueError("batch was packed for a different bond vocabulary")
        st = _stream()
        if not self._tables_valid or getattr(self, "_packed6_gen", None) not in (None, self.planned_gen()):
            self.refresh_tables()
        if getattr(batch, "is_compact", False):
            if keep or unfused_messages or not self.use_fused(batch) or not self.compact_supported():
                raise _lib.ImpError("a compact-feed batch can only run through the default fused half-precision kernel")
            return self._forward_fused(batch, None, st)
        g = batch.c_struct()
        N, P = batch.n_atoms, batch.n_pairs
        if self.wide_supported():
            if keep or unfused_messages:
                raise _lib.ImpError("the wide tensor path keeps no per-layer intermediates (use precision='fp32')")
            return self._forward_wide(batch, g, st)
        if not keep and not unfused_messages and self.use_fused(batch):
            return self._forward_fused(batch, g, st)
        inter = {}
        if keep:
            h = [torch.empty(N * d, dtype=torch.float32, device=self.device) for _ in range(S + 1)]
            aggs = [torch.empty(N * d, dtype=torch.float32, device=self.device) for _ in range(S)]
        else:
            hb = [self._buf("h0", N * d), self._buf("h1", N * d)]
            h = [hb[i % 2] for i in range(S + 1)]
            aggs = [self._buf("agg", N * d)] * S
        folded = (not keep and not unfused_messages and self.precision != "fp32" and d == 32 and "bucket_perm" in batch.dev
                  and not getattr(self, "simt_messages", False))
        io16 = folded and not getattr(self, "fp32_messages", False)
        planned = io16 and s["bond_vocab_size"] <= 256 and not (self.tc_flags() & _lib.TC_MSG_ONE_CHUNK_PER_CTA)
        if planned:  # per-batch index plan of the grouped message kernel (chunk offsets, bucket-ordered src / bond|mult)
            plan = self._buf("msg_plan", _lib.load().imp_edge_messages_tc16_plan_bytes(batch.n_unique, s["bond_vocab_size"]),
                             torch.uint8)
            _lib.call("imp_edge_messages_tc16_plan", C.byref(g), plan.data_ptr(), st)
        if io16:
            h16 = [self._buf("h16_0", N * d, torch.int16), self._buf("h16_1", N * d, torch.int16)]
            _lib.call("imp_embed_atoms16", self._ptr("atom_emb"), s["atom_vocab_size"], batch.dev["atom_id"].data_ptr(), N, d,
                      self.tc_flags(), h[0].data_ptr(), h16[0].data_ptr(), st)
        else:
            _lib.call("imp_embed_atoms", self._ptr("atom_emb"), s["atom_vocab_size"], batch.dev["atom_id"].data_ptr(), N, d,
                      h[0].data_ptr(), st)
        for i in range(S):
            if folded:
                # tensor staged path without intermediates: grouped tcgen05 message GEMM, then Reduce folded into the
                # load stage of the tcgen05 GatedUpdate (agg is never written); by default 16-bit rows in and out of the
                # message kernel (operand-format copy of h, operand-format messages)
                mbase, gbase = self._ws["msg_packed"].data_ptr(), self._ws["gru_packed"].data_ptr()
                cws = self._buf("msg_chunks", 2 * s["bond_vocab_size"] + 1, torch.int32)
                pm = (mbase + self._msg_pack_bytes * i, mbase + self._msg_pack_bytes * (S + i))
                pg = (gbase + self._gru_pack_bytes * i, gbase + self._gru_pack_bytes * (S + i))
                if io16:
                    msg16 = self._buf("msg16", batch.n_unique * d, torch.int16)
                    if planned:
                        _lib.call("imp_edge_messages_tc16_planned", C.byref(g), plan.data_ptr(), h16[i % 2].data_ptr(), d, pm[0],
                                  pm[1], self.tc_flags(), msg16.data_ptr(), st)
                    else:
                        _lib.call("imp_edge_messages_tc16", C.byref(g), h16[i % 2].data_ptr(), d, pm[0], pm[1], self.tc_flags(),
                                  msg16.data_ptr(), cws.data_ptr(), st)
                    _lib.call("imp_reduce_gated_update_tc16", C.byref(g), h[i].data_ptr(), msg16.data_ptr(), d, pg[0], pg[1],
                              C.c_float(self.LN_EPS), self.tc_flags(), h[i + 1].data_ptr(), h16[(i + 1) % 2].data_ptr(), st)
                else:
                    msg = self._buf("msg", batch.n_unique * d)
                    _lib.call("imp_edge_messages_tc", C.byref(g), h[i].data_ptr(), d, pm[0], pm[1], self.tc_flags(),
                              msg.data_ptr(), cws.data_ptr(), st)
                    _lib.call("imp_reduce_gated_update_tc", C.byref(g), h[i].data_ptr(), msg.data_ptr(), d, pg[0], pg[1],
                              C.c_float(self.LN_EPS), self.tc_flags(), h[i + 1].data_ptr(), st)
                continue
            if unfused_messages:
                msg = self._buf("msg", batch.n_unique * d)
                _lib.call("imp_edge_messages", C.byref(g), h[i].data_ptr(), d, self.table_ptr(0, i, False),
                          self.table_ptr(1, i, False), msg.data_ptr(), st)
                _lib.call("imp_segment_sum", C.byref(g), msg.data_ptr(), d, aggs[i].data_ptr(), st)
            elif (self.precision != "fp32" and d == 32 and not getattr(self, "simt_messages", False)
                  and "bucket_perm" in batch.dev):  # device-packed batches carry no bond buckets: CSR-order kernel
                # bond-type-grouped tcgen05 GEMM over gathered source rows, then the CSR segment sum (csrc/msg_tc.cu)
                mbase = self._ws["msg_packed"].data_ptr()
                msg = self._buf("msg", batch.n_unique * d)
                cws = self._buf("msg_chunks", 2 * s["bond_vocab_size"] + 1, torch.int32)
                _lib.call("imp_edge_messages_tc", C.byref(g), h[i].data_ptr(), d, mbase + self._msg_pack_bytes * i,
                          mbase + self._msg_pack_bytes * (S + i), self.tc_flags(), msg.data_ptr(), cws.data_ptr(), st)
                _lib.call("imp_segment_sum", C.byref(g), msg.data_ptr(), d, aggs[i].data_ptr(), st)
            elif d == 32 and "bucket_perm" in batch.dev and not getattr(self, "simt_messages", False):
                msg = self._buf("msg", batch.n_unique * d)
                if getattr(self, "fp32_tensor", False) and s["bond_vocab_size"] <= 256:
                    # fp32-class on the tensor cores (3xTF32, csrc/msg_tc32.cu), from the per-batch index plan
                    plan32 = getattr(batch, "_msg_plan32", None)
                    if plan32 is None:
                        nbp = _lib.load().imp_edge_messages_tc16_plan_bytes(batch.n_unique, s["bond_vocab_size"])
                        plan32 = batch._msg_plan32 = torch.empty(max(int(nbp), 16), dtype=torch.uint8, device=self.device)
                        _lib.call("imp_edge_messages_tc16_plan", C.byref(g), plan32.data_ptr(), st)
                    _lib.call("imp_edge_messages_grouped_tc32_planned", C.byref(g), plan32.data_ptr(), h[i].data_ptr(), d,
                              self.table_ptr(0, i, False), self.table_ptr(1, i, False), 2, msg.data_ptr(), st)  # 2: rounded splits
                else:
                    # exact fp32, bucket-grouped: T[b] staged once per chunk of 128 entries, then the CSR segment sum
                    cws = self._buf("msg_chunks", 2 * s["bond_vocab_size"] + 1, torch.int32)
                    _lib.call("imp_edge_messages_grouped", C.byref(g), h[i].data_ptr(), d, self.table_ptr(0, i, False),
                              self.table_ptr(1, i, False), 0, msg.data_ptr(), cws.data_ptr(), st)
                _lib.call("imp_segment_sum", C.byref(g), msg.data_ptr(), d, aggs[i].data_ptr(), st)
            else:
                _lib.call("imp_message_agg", C.byref(g), h[i].data_ptr(), d, self.table_ptr(0, i, True),
                          self.table_ptr(1, i, True), aggs[i].data_ptr(), st)
            if self.precision == "fp32" and d > 64:  # wide atom states: tiled-GEMM GatedUpdate
                wc, wa = self._gru_struct("cat", i), self._gru_struct("an", i)
                ws = self._buf("gru_wide_ws", 3 * N * d)
                _lib.call("imp_gated_update_wide", h[i].data_ptr(), aggs[i].data_ptr(), N, batch.n_cat_atoms, d, C.byref(wc),
                          C.byref(wa), C.c_float(self.LN_EPS), h[i + 1].data_ptr(), ws.data_ptr(), st)
            elif self.precision == "fp32":
                wc, wa = self._gru_struct("cat", i), self._gru_struct("an", i)
                if getattr(self, "fp32_tensor", False) and d == 32:  # 3xTF32 tensor-core GatedUpdate (csrc/fwd_tc32.cu)
                    _lib.call("imp_gated_update_tc32", h[i].data_ptr(), aggs[i].data_ptr(), N, batch.n_cat_atoms, d, C.byref(wc),
                              C.byref(wa), C.c_float(self.LN_EPS), h[i + 1].data_ptr(), None, None, None, st)
                else:
                    _lib.call("imp_gated_update", h[i].data_ptr(), aggs[i].data_ptr(), N, batch.n_cat_atoms, d, C.byref(wc),
                              C.byref(wa), C.c_float(self.LN_EPS), h[i + 1].data_ptr(), st)
            else:
                base = self._ws["gru_packed"].data_ptr()
                _lib.call("imp_gated_update_tc", h[i].data_ptr(), aggs[i].data_ptr(), N, batch.n_cat_atoms, d,
                          base + self._gru_pack_bytes * i, base + self._gru_pack_bytes * (S + i), C.c_float(self.LN_EPS),
                          self.tc_flags(), h[i + 1].data_ptr(), st)
        out = torch.empty(P, dtype=torch.float32, device=self.device)
        fp, mix = s["fp_size"], s["mixing_size"]
        visc = s["kind"] == "viscosity"
        aux = None
        if keep:
            aux = torch.empty(P * (2 * d + 2 * fp + mix + (3 if visc else 0)), dtype=torch.float32, device=self.device)
        rc, ra = self._readout_struct("cat"), self._readout_struct("an")
        auxp = C.c_void_p(aux.data_ptr()) if keep else None
        if visc and batch.dev_T is None:
            raise ValueError("viscosity model needs batch.temperature")
        if self.precision != "fp32" and not keep and d <= 32 and fp <= 32 and mix <= 32:
            # 16-bit operand routes: GlobalSumPool (one warp per molecule, whole rows) + the fp32 one-thread-per-pair readout of the
            # fused path.  (imp_pool_head_* is the readout of the 1e-5 path: one warp per pair, double accumulation -- 0.30
            # instead of 0.18 ms per 64 k pairs.)
            pooled = self._buf("pooled", 2 * P * d)
            _lib.call("imp_global_sum_pool", batch.dev["mol_ptr"].data_ptr(), batch.dev["atom_id"].data_ptr(), 2 * P,
                      h[S].data_ptr(), d, pooled.data_ptr(), st)
            if visc:
                _lib.call("imp_readout_visc", pooled.data_ptr(), P, d, fp, mix, C.byref(rc), C.byref(ra),
                          self._ptr("head.kernel"), self._ptr("head.bias"), batch.dev_T.data_ptr(), out.data_ptr(), None, st)
            else:
                _lib.call("imp_readout_mp", pooled.data_ptr(), P, d, fp, mix, fp, C.byref(rc), C.byref(ra),
                          self._ptr("head1.kernel"), self._ptr("head1.bias"), self._ptr("head2.kernel"),
                          self._ptr("head2.bias"), out.data_ptr(), None, st)
        elif visc:
            _lib.call("imp_pool_head_visc", C.byref(g), h[S].data_ptr(), d, fp, mix, C.byref(rc), C.byref(ra),
                      self._ptr("head.kernel"), self._ptr("head.bias"), batch.dev_T.data_ptr(), out.data_ptr(), auxp, st)
        else:
            _lib.call("imp_pool_head_mp", C.byref(g), h[S].data_ptr(), d, fp, mix, fp, C.byref(rc), C.byref(ra),
                      self._ptr("head1.kernel"), self._ptr("head1.bias"), self._ptr("head2.kernel"),
                      self._ptr("head2.bias"), out.data_ptr(), auxp, st)
        if keep:
            inter = {"h": [t.view(N, d) for t in h], "agg": [t.view(N, d) for t in aggs],
                     "aux": aux.view(P, -1)}
        return (out, inter) if keep else out

    def _forward_fused(self, batch, g, st):
        """imp_mpnn_forward_fused (embed + all steps + pool, one kernel) then the readout kernel."""
        import torch

        s = self.spec
        d, P = s["atom_dim"], batch.n_pairs
        fp, mix = s["fp_size"], s["mixing_size"]
        pooled = self._fused_pooled(batch, g, st)
        out = torch.empty(P, dtype=torch.float32, device=self.device)
        rc, ra = self._readout_struct("cat"), self._readout_struct("an")
        if s["kind"] == "viscosity":
            if batch.dev_T is None:
                raise ValueError("viscosity model needs batch.temperature")
            _lib.call("imp_readout_visc", pooled.data_ptr(), P, d, fp, mix, C.byref(rc), C.byref(ra),
                      self._ptr("head.kernel"), self._ptr("head.bias"), batch.dev_T.data_ptr(), out.data_ptr(), None, st)
        else:
            _lib.call("imp_readout_mp", pooled.data_ptr(), P, d, fp, mix, fp, C.byref(rc), C.byref(ra),
                      self._ptr("head1.kernel"), self._ptr("head1.bias"), self._ptr("head2.kernel"),
                      self._ptr("head2.bias"), out.data_ptr(), None, st)
        return out

    def pooled_sums(self, batch):
        """GlobalSumPool outputs of both towers, [2P, d] tower-major (cations first), through the forward path this model is
        configured for: the fused kernel when the shape allows it, else the staged kernels.  (The transfer-learning head,
        ionic_mpnn_b200/transfer.py, starts from these.)"""
        import torch

        if batch.dev is None:
            batch.to(self.device)
        if not self._tables_valid or getattr(self, "_packed6_gen", None) not in (None, self.planned_gen()):
            self.refresh_tables()
        st = _stream()
        if self.use_fused(batch) and not self.wide_supported():
            compact = getattr(batch, "is_compact", False)
            return self._fused_pooled(batch, None if compact else batch.c_struct(), st)
        if batch.dev_T is None and self.spec["kind"] == "viscosity":  # the staged readout wants a temperature; any value does
            batch.dev_T = torch.full((batch.n_pairs,), 300.0, dtype=torch.float32, device=self.device)
        _, inter = self.forward_packed(batch, keep=True)
        d = self.spec["atom_dim"]
        aux = inter["aux"]
        return torch.cat([aux[:, :d], aux[:, d:2 * d]], dim=0).contiguous()

    def _fused_pooled(self, batch, g, st):
        """The fused forward up to the molecule sums: returns the [2P, d] workspace tensor."""
        import torch

        s = self.spec
        d, S, P = s["atom_dim"], s["num_steps"], batch.n_pairs
        pooled = self._buf("pooled", 2 * P * d)
        status = self._ws.get("status")
        if status is None:
            status = self._ws["status"] = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.planned_supported(batch):
            # tile plan of this batch (integer-only, rebuilt per call: it depends on the batch alone), then the planned kernel
            cg = batch.compact_struct() if g is None else None
            nb = _lib.load().imp_fused_plan_bytes(P, batch.n_atoms, batch.n_unique, batch.max_mol_atoms)
            if nb < 0:
                raise _lib.ImpError(f"imp_fused_plan_bytes: {nb}")
            # imp_fused_plan_bytes is the bound that can never overflow (one tile per molecule in the worst case: 4 KB per
            # pair).  Best-fit tiles are ~98 % full, so the buffer is sized for plan_slack x (atoms of the larger tower /
            # 128) tiles per tower instead; a batch that needs more (many rows closed by the entry limit) reports status 2
            # and check_status() switches this model to the safe bound.
            slack = getattr(self, "plan_slack", "auto")
            if slack is not None:
                tower = max(batch.n_cat_atoms, batch.n_atoms - batch.n_cat_atoms)
                if slack == "auto":  # well-filled tiles (1.25x), but never less than the row-limit bound: a tile closed by the
                    # row limit has at least 129 - max_mol_atoms rows in use
                    slack = max(1.25, 128.0 / max(129 - int(batch.max_mol_atoms), 1))
                cap = max(int(slack * tower / 128), (P + 31) // 32) + (P + 255) // 256 + 16  # (at most 32 molecules per tile)
                nb = min(nb, 256 + 2 * cap * 2048)
            plan = self._buf("fused_plan", nb, torch.uint8)
            if getattr(batch, "is_narrow", False):  # compact feed with 16-bit entry words
                _lib.call("imp_fused_plan_compact16", C.byref(cg), s["atom_vocab_size"], batch.max_mol_atoms, plan.data_ptr(), nb, st)
            else:
                _lib.call("imp_fused_plan", C.byref(g) if g is not None else None, C.byref(cg) if cg is not None else None,
                          s["atom_vocab_size"], batch.max_mol_atoms, plan.data_ptr(), nb, st)
            _lib.call("imp_mpnn_forward_fused_planned", plan.data_ptr(), P, batch.n_atoms, batch.n_cat_atoms, batch.bond_vocab,
                      self._ptr("atom_emb"), s["atom_vocab_size"], self._ptr("bond_emb"), d, s["bond_dim"], S,
                      self._ws["fused_packed" if self.tc_flags() & _lib.TC_GEN5 else "fused_packed6"].data_ptr(),
                      C.c_float(self.LN_EPS), self.tc_flags() | {7: _lib.TC_GEN7, 8: _lib.TC_GEN8}.get(self.planned_gen(), 0),
                      pooled.data_ptr(), st)
        elif g is None:
            if getattr(batch, "is_narrow", False):
                raise _lib.ImpError("the narrow compact feed (16-bit entry words) is read by the planned forward only")
            cg = batch.compact_struct()
            _lib.call("imp_mpnn_forward_fused_compact", C.byref(cg), self._ptr("atom_emb"), s["atom_vocab_size"],
                      self._ptr("bond_emb"), d, s["bond_dim"], S, self._ws["fused_packed"].data_ptr(), C.c_float(self.LN_EPS),
                      self.tc_flags() & ~_lib.TC_GEN5, batch.max_mol_atoms, pooled.data_ptr(), status.data_ptr(), st)
        else:
            _lib.call("imp_mpnn_forward_fused", C.byref(g), self._ptr("atom_emb"), s["atom_vocab_size"], self._ptr("bond_emb"),
                      d, s["bond_dim"], S, self._ws["fused_packed"].data_ptr(), C.c_float(self.LN_EPS),
                      self.tc_flags() & ~_lib.TC_GEN5, batch.max_mol_atoms, pooled.data_ptr(), status.data_ptr(), st)
        return pooled

    def _forward_wide(self, batch, g, st):
        """imp_mpnn_forward_wide (embed, 3 tcgen05 GEMM kernels per step, pool) then the readout kernel."""
        import torch

        s = self.spec
        d, S, P = s["atom_dim"], s["num_steps"], batch.n_pairs
        pooled = self._buf("pooled", 2 * P * d)
        wsb = _lib.load().imp_wide_workspace_bytes(batch.n_atoms, d)
        ws = self._buf("wide_ws", wsb, torch.uint8)
        if getattr(self, "wide_per_stage_calls", False):  # the same launches as separate ABI calls (bench: per-kernel timing)
            wb = _lib.load().imp_wide_pack_bytes(d, s["bond_dim"])
            pk, f, w = self._ws["wide_packed"].data_ptr(), self.tc_flags(), ws.data_ptr()
            _lib.call("imp_wide_embed", C.byref(g), self._ptr("atom_emb"), s["atom_vocab_size"], d, w, st)
            for i in range(S):
                pc, pa = pk + wb * i, pk + wb * (S + i)
                _lib.call("imp_wide_message", C.byref(g), self._ptr("bond_emb"), d, s["bond_dim"], pc, pa, f, w, st)
                if f & _lib.TC_WIDE_SPLIT_GRU:
                    _lib.call("imp_wide_gates", C.byref(g), d, pc, pa, f, w, st)
                    _lib.call("imp_wide_candidate", C.byref(g), d, pc, pa, C.c_float(self.LN_EPS), f, w, st)
                else:
                    _lib.call("imp_wide_gated_update", C.byref(g), d, pc, pa, C.c_float(self.LN_EPS), f, w, st)
            _lib.call("imp_wide_pool", C.byref(g), d, w, pooled.data_ptr(), st)
        else:
            _lib.call("imp_mpnn_forward_wide", C.byref(g), self._ptr("atom_emb"), s["atom_vocab_size"], self._ptr("bond_emb"),
                      d, s["bond_dim"], S, self._ws["wide_packed"].data_ptr(), C.c_float(self.LN_EPS), self.tc_flags(),
                      ws.data_ptr(), pooled.data_ptr(), st)
        return self._readout(batch, pooled, st)

    def _readout(self, batch, pooled, st):
        import torch

        s = self.spec
        d, P, fp, mix = s["atom_dim"], batch.n_pairs, s["fp_size"], s["mixing_size"]
        out = torch.empty(P, dtype=torch.float32, device=self.device)
        rc, ra = self._readout_struct("cat"), self._readout_struct("an")
        if s["kind"] == "viscosity":
            if batch.dev_T is None:
                raise ValueError("viscosity model needs batch.temperature")
            _lib.call("imp_readout_visc", pooled.data_ptr(), P, d, fp, mix, C.byref(rc), C.byref(ra),
                      self._ptr("head.kernel"), self._ptr("head.bias"), batch.dev_T.data_ptr(), out.data_ptr(), None, st)
        else:
            _lib.call("imp_readout_mp", pooled.data_ptr(), P, d, fp, mix, fp, C.byref(rc), C.byref(ra),
                      self._ptr("head1.kernel"), self._ptr("head1.bias"), self._ptr("head2.kernel"),
                      self._ptr("head2.bias"), out.data_ptr(), None, st)
        return out

    def launches_per_forward(self, batch=None):
        """Kernels enqueued by forward_packed (tables / packs already valid)."""
        S = self.spec["num_steps"]
        if self.wide_supported():
            return 1 + 2 * S + 1 + 1  # embed, (messages, GatedUpdate) per step, pool, readout
        if batch is not None and self.use_fused(batch):
            return 4 if self.planned_supported(batch) else 2  # [plan header, tile plan,] fused forward, readout
        grouped = batch is None or "bucket_perm" in (batch.dev or {})
        if self.precision != "fp32" and self.spec["atom_dim"] == 32 and grouped:
            return 3 + 2 * S + 2      # embed, message plan (2), (grouped message GEMM, Reduce + GatedUpdate) per step, pool, readout
        if self.spec["atom_dim"] == 32 and grouped and getattr(self, "fp32_tensor", False):
            return 1 + 3 * S + 1      # embed, (planned 3xTF32 messages, segment sum, 3xTF32 GatedUpdate) per step, pool + head
        if self.spec["atom_dim"] == 32 and grouped:
            return 1 + 4 * S + 1      # embed, (chunk scan, grouped messages, segment sum, GatedUpdate) per step, pool + head
        return 1 + 2 * S + 1          # embed, (CSR-order messages, GatedUpdate) per step, pool + head

    # -- Keras-like surface ---------------------------------------------------------------------
    def pack(self, x):
        if isinstance(x, PackedGraphBatch):
            return x
        if isinstance(x, dict):
            return pack_padded(x, self.spec["bond_vocab_size"])
        return pack_records(x, self.spec["bond_vocab_size"])

    def predict(self, x, batch_size=None, verbose=0):
        """``x``: the reference's padded input dict (train_viscosity.py:306-314), a list of records, or a
        PackedGraphBatch.  Returns numpy ``(P, 1)`` like Keras.  ``batch_size`` is accepted for signature
        compatibility; the packed batch runs in one pass."""
        import torch

        batch = self.pack(x)
        if "status" in self._ws:
            self._ws["status"].zero_()
        out = self.forward_packed(batch)
        torch.cuda.current_stream().synchronize()
        try:
            self.check_status()
        except _lib.PlanCapacityError:  # the plan buffer was sized by plan_slack and this batch needed more: once more
            out = self.forward_packed(batch)
            torch.cuda.current_stream().synchronize()
            self.check_status()
        return out.cpu().numpy().reshape(-1, 1)

    __call__ = predict

    def check_status(self):
        """Reads (and clears) the fused kernel's status word: non-zero means a molecule did not fit a 128-row tile, i.e. the
        batch's ``max_mol_atoms`` was wrong and the predictions of that launch are invalid.  ``predict`` calls it after its
        synchronisation; callers of the enqueue-only entry points (``forward_packed``, ``predict_stream``) call it once they
        have synchronised."""
        import torch

        plan = self._ws.get("fused_plan")
        if plan is not None and plan.numel() >= 20:
            code = int(plan[16:20].view(torch.int32).item())
            if code == 2:
                plan[16:20].zero_()
                self.plan_slack = None
                raise _lib.PlanCapacityError("fused forward: the tile plan needed more tiles than plan_slack allowed; the model "
                                             "now sizes the plan buffer by the safe bound -- run the batch again")
            if code != 0:
                plan[16:20].zero_()
                raise _lib.ImpError("fused forward: the tile plan refused the batch (a molecule with > 128 atoms, a row with "
                                    "> 31 entries or > 336 entries per molecule); set model.use_plan = False (self-contained "
                                    "fused kernel) or fused=False")
        st = self._ws.get("status")
        if st is None or st.numel() != 1:
            return
        if int(st.item()) != 0:
            st.zero_()
            raise _lib.ImpError("fused forward: a molecule does not fit one 128-row tile (max_mol_atoms of the batch was wrong); "
                                "use fused=False / 'auto' with a correct max_mol_atoms")

    def predict_stream(self, chunks, out=None, compact="auto"):
        """Pipelined prediction over a list of host-resident packed chunks (the cfg-3 "inference sweep" shape): the
        H2D copy of chunk i+1 runs on a copy stream while chunk i computes; predictions are copied back into ``out``
        (a pinned float32 tensor of the total pair count, allocated if None).  Two device staging slots.  With
        ``compact`` ('auto': when the fused kernel runs and the batch fits) the compact input feed is copied instead of
        the int32 CSR arrays: 0.49 instead of 1.15 KB per pair over PCIe, bit-identical predictions.
        Returns (out, bytes_h2d).  Nothing is synchronised on return: the caller syncs the current stream."""
        import torch

        from .graph import COMPACT16_FIELDS, COMPACT_FIELDS, FUSED_FIELDS, GRAPH_FIELDS, DeviceSlot

        total = sum(c.n_pairs for c in chunks)
        if out is None:
            out = torch.empty(total, dtype=torch.float32).pin_memory()
        if not self._tables_valid or getattr(self, "_packed6_gen", None) not in (None, self.planned_gen()):
            self.refresh_tables()
        st = getattr(self, "_stream_state", None)
        if st is None:
            st = self._stream_state = {"copy": torch.cuda.Stream(device=self.device), "slots": None, "fields": None}
        compute = torch.cuda.current_stream()
        copy = st["copy"]
        nbytes = 0
        off = 0
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        # the two staging slots and the events of their last readers persist across calls: the first copies of a sweep's
        # next call run under the tail of the previous call's kernels instead of waiting for the whole compute stream
        done = st.setdefault("done", [None, None])
        for i, ch in enumerate(chunks):
            fused = self.use_fused(ch)
            use_compact = fused and self.compact_supported() and compact in ("auto", True)
            if use_compact:
                try:
                    ch.pin_compact()
                except _lib.ImpError:
                    if compact is True:
                        raise
                    use_compact = False
            if not use_compact:
                ch.pin()
            narrow = use_compact and self.planned_supported(ch) and getattr(ch, "narrow_ok", False)
            fields = (COMPACT16_FIELDS if narrow else COMPACT_FIELDS) if use_compact else (FUSED_FIELDS if fused else GRAPH_FIELDS)
            if st["slots"] is None or st["fields"] != fields:
                copy.wait_stream(compute)  # the old slots may still be read
                st["slots"] = [DeviceSlot(self.device, fields), DeviceSlot(self.device, fields)]
                st["fields"] = fields
                done[0] = done[1] = None
            slot = st["slots"][i % 2]
            if done[i % 2] is not None:
                copy.wait_event(done[i % 2])  # the kernels that read this slot two chunks ago have finished
            else:
                copy.wait_stream(compute)
            nbytes += slot.load(ch, copy)
            ready[i % 2] = torch.cuda.Event()
            ready[i % 2].record(copy)
            compute.wait_event(ready[i % 2])
            o = self.forward_packed(slot)
            out[off:off + ch.n_pairs].copy_(o, non_blocking=True)
            off += ch.n_pairs
            done[i % 2] = torch.cuda.Event()
            done[i % 2].record(compute)
        return out, nbytes


def load_model(path, precision="fp32", **kw):
    """``tf.keras.models.load_model(path)`` for the reference's archives (train_melting_point_transfer.py:78-93): builds the
    model whose spec the archive's variable shapes imply and loads them."""
    from . import keras_io

    config, data = keras_io.read_keras(path)
    kind, params, _ = keras_io.params_from_keras(config, data)
    m = MPNNModel(keras_io.spec_from_params(kind, params), precision=precision, **kw)
    if kind == "transfer":
        cur = m.get_weights()
        params = {**{k: cur[k] for k in cur if k.startswith("head")}, **{k: v for k, v in params.items() if not k.startswith("head")}}
    m.set_weights(params)
    return m


def make_spec(kind="viscosity", atom_vocab_size=124, bond_vocab_size=72, atom_dim=32, bond_dim=8, fp_size=32,
              mixing_size=20, num_steps=4):
    if kind == "melting_point":
        bond_dim = atom_dim * atom_dim  # train_melting_point.py:146
    return dict(kind=kind, atom_vocab_size=atom_vocab_size, bond_vocab_size=bond_vocab_size, atom_dim=atom_dim,
                bond_dim=bond_dim, fp_size=fp_size, mixing_size=mixing_size, num_steps=num_steps)
