"""Training step of the reference (train_viscosity.py:227-230,328-338; train_melting_point.py:210-214) on the B200
kernels: loss = mean((y - yhat)^2) + l2 kernel regularisers, per-variable clip_by_norm(1.0), Adam(1e-3).

Forward = the staged fp32 kernels with every h_i / agg_i kept; backward = csrc/bwd_fp32.cu (B1..B6); the flat gradient
bucket is summed over ranks with ONE all-reduce (torch.distributed, NCCL) when a process group is initialised --
SURVEY 8e; the only collective of the whole framework.  The bucket's four-float tail carries the squared-error sum, the
pair count and the two per-occurrence Embedding norms through the same collective.  Mixed into MPNNModel (model.py).

Clip semantics [Keras 2.12, restated in oracle/ref_model.py:adam_step]: every variable is clipped by its own norm; for
the two Embedding variables that norm is taken over the per-occurrence gradient rows (the optimizer clips the
IndexedSlices gradient before de-duplicating it), computed by imp_sumsq / imp_bond_occurrence_norm2
(``embedding_clip="occurrence"``, the default; ``"dense"`` clips by the dense gradient's norm)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

TOWERS = ("cat", "an")
DT_CHUNK = 2048  # entries per CTA of the dTable partial kernel


def l2_terms(spec):
    """(variable, coefficient) of the kernel regularisers: train_viscosity.py:189; train_melting_point.py:172,196."""
    if spec["kind"] == "viscosity":
        return {"cat_fp.kernel": 1e-4, "an_fp.kernel": 1e-4}
    return {"cat_fp.kernel": 1e-5, "an_fp.kernel": 1e-5, "head1.kernel": 1e-5}


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def prepare_training(batch, device):
    """Per-batch index structures of the backward pass, built once on the host and uploaded:
    entry_dst[e] (destination atom of CSR entry e) and the split of the (tower, bond) buckets into chunks."""
    import torch

    if getattr(batch, "train_dev", None) is not None:
        return batch.train_dev
    rp = batch.host["row_ptr"].astype(np.int64)
    entry_dst = np.repeat(np.arange(batch.n_atoms, dtype=np.int32), np.diff(rp)).astype(np.int32)
    bp = batch.host["bucket_ptr"].astype(np.int64)
    nb = len(bp) - 1
    sizes = np.diff(bp)
    n_chunks_per = (sizes + DT_CHUNK - 1) // DT_CHUNK
    bcp = np.zeros(nb + 1, np.int32)
    bcp[1:] = np.cumsum(n_chunks_per)
    total = int(bcp[-1])
    which = np.repeat(np.arange(nb), n_chunks_per)
    idx_in = np.arange(total) - bcp[which]
    begin = (bp[which] + idx_in * DT_CHUNK).astype(np.int32)
    end = np.minimum(bp[which + 1], begin.astype(np.int64) + DT_CHUNK).astype(np.int32)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in
           dict(entry_dst=entry_dst if len(entry_dst) else np.zeros(1, np.int32), chunk_begin=begin if total else np.zeros(1, np.int32),
                chunk_end=end if total else np.zeros(1, np.int32), bucket_chunk_ptr=bcp).items()}
    dev["n_chunks"] = total
    dev["n_cat_unique"] = int(rp[batch.n_cat_atoms])
    if getattr(batch, "symmetric", None) is None:
        batch.symmetric = entries_are_symmetric(batch)
    batch.train_dev = dev
    return dev


def entries_are_symmetric(batch):
    """True when the live entry multiset is closed under reversal: every (dst <- src, bond, mult) has its mirror
    (src <- dst, bond, mult).  The message backward (dh += T[b]^T dagg[src] over the SAME CSR) is the true transpose only
    then.  Holds by construction for reverse-doubled batches (train_viscosity.py:87-91) without ``max_edges`` truncation;
    checked on the host (one sort) for everything else."""
    rp = batch.host["row_ptr"].astype(np.int64)
    if batch.n_unique == 0:
        return True
    dst = np.repeat(np.arange(batch.n_atoms, dtype=np.int64), np.diff(rp))
    src = batch.host["col_src"].astype(np.int64)
    bm = batch.host["edge_bm"].astype(np.int64) & 0xFFFFFFFF
    n = np.int64(batch.n_atoms)
    fwd = (dst * n + src)
    rev = (src * n + dst)
    o1, o2 = np.lexsort((bm, fwd)), np.lexsort((bm, rev))
    return bool(np.array_equal(fwd[o1], rev[o2]) and np.array_equal(bm[o1], bm[o2]))


class TrainMixin:
    """loss_and_grads / train_step for MPNNModel."""

    def _train_state(self):
        import torch

        st = getattr(self, "_train", None)
        if st is not None:
            return st
        s = self.spec
        if s["atom_dim"] != 32:
            raise _lib.ImpError("the backward kernels are built for atom_dim 32")
        n = self.flat.numel()
        # gradient bucket + tail [sse, pair count, occurrence norm^2 of atom_emb, of bond_emb]: one collective carries all
        st = {"grad": torch.zeros(n + 4, dtype=torch.float32, device=self.device),
              "m": torch.zeros(n, dtype=torch.float32, device=self.device),
              "v": torch.zeros(n, dtype=torch.float32, device=self.device), "step": 0}
        shapes = {k: tuple(self.params[k].shape) for k in self.var_names}
        st["g"] = {k: st["grad"][self.var_off[k]: self.var_off[k] + int(np.prod(shapes[k]))].view(*shapes[k])
                   for k in self.var_names}
        offs = [self.var_off[k] for k in self.var_names] + [n]
        st["var_off"] = torch.tensor(offs, dtype=torch.int64, device=self.device)
        l2 = l2_terms(s)
        st["var_l2"] = torch.tensor([l2.get(k, 0.0) for k in self.var_names], dtype=torch.float32, device=self.device)
        st["norms2"] = torch.zeros(2 * len(self.var_names), dtype=torch.float32, device=self.device)
        st["tail"] = st["grad"][n:n + 4]
        st["sse"] = st["grad"][n:n + 1]
        st["occ_var"] = torch.tensor([self.var_names.index("atom_emb"), self.var_names.index("bond_emb")], dtype=torch.int32,
                                     device=self.device)
        st["loss"] = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._train = st
        return st

    def bond_occurrence_supported(self):
        """imp_bond_occurrence_norm2 is built for the viscosity shape (bond_dim 8).  For the melting-point model
        (bond_dim = 1024) the per-occurrence norm of the bond Embedding costs ~1 MFLOP per edge entry and step -- the
        arithmetic the reference itself does -- and is not built: its bond Embedding is clipped by the dense norm
        ("Embedding clip norms")."""
        return self.spec["atom_dim"] == 32 and self.spec["bond_dim"] == 8

    def _forward_kept(self, batch):
        """The staged fp32 forward up to the molecule sums with every h_i / agg_i kept (the tape of the backward pass).
        Returns a dict: h, agg (lists of device tensors), pooled / dpooled [2P, d], the bond tables and chunk workspace."""
        import torch

        s = self.spec
        d, S, K, Vb = s["atom_dim"], s["num_steps"], s["bond_dim"], s["bond_vocab_size"]
        st = self._train_state()
        if batch.dev is None:
            batch.to(self.device)
        if batch.dev_y is None:
            raise ValueError("training needs batch.target")
        tr = prepare_training(batch, self.device)
        if not batch.symmetric:
            raise _lib.ImpError("training needs a batch whose live entries are symmetric (every dst<-src entry has its src<-dst "
                                "mirror with the same bond and multiplicity): the message backward runs over the forward CSR. "
                                "pack_records / pack_flat(double_edges=True) without max_edges produce such batches.")
        g = batch.c_struct()
        N, P = batch.n_atoms, batch.n_pairs
        sm = _stream()
        G = st["g"]
        lib = _lib.load()
        # ---- bond-matrix tables for this weight set: row-major [V_b, d, d] (the bucket-grouped message kernel reads T[b]
        # for the forward and, transposed on the fly, for the backward)
        n = 2 * S
        per = Vb * d * d
        tab = self._buf("tr_table", n * per)
        W = (C.c_void_p * n)(*[self.params[f"{t}_bmm_{i}.bond_transform"].data_ptr() for t in TOWERS for i in range(S)])
        T1 = (C.c_void_p * n)(*[tab.data_ptr() + 4 * per * j for j in range(n)])
        _lib.call("imp_bond_table", self._ptr("bond_emb"), Vb, K, d, n, W, T1, None, sm)
        msg = self._buf("tr_msg", batch.n_unique * d)
        cws = self._buf("tr_msg_chunks", 2 * Vb + 1, torch.int32)
        # ---- forward, everything kept
        h = [self._buf(f"tr_h{i}", N * d) for i in range(S + 1)]
        agg = [self._buf(f"tr_agg{i}", N * d) for i in range(S)]
        gates = ([[self._buf(f"tr_{n}{i}", N * d) for n in ("z", "r", "ht")] for i in range(S)]
                 if getattr(self, "store_gates", True) else None)
        _lib.call("imp_embed_atoms", self._ptr("atom_emb"), s["atom_vocab_size"], batch.dev["atom_id"].data_ptr(), N, d,
                  h[0].data_ptr(), sm)
        tc_msg = d == 32 and Vb <= 256 and getattr(self, "tc_forward", True)
        if tc_msg and tr.get("msg_plan") is None:  # per-batch index plan of the tensor-core message kernel (built once)
            nbp = lib.imp_edge_messages_tc16_plan_bytes(batch.n_unique, Vb)
            tr["msg_plan"] = torch.empty(max(int(nbp), 16), dtype=torch.uint8, device=self.device)
            _lib.call("imp_edge_messages_tc16_plan", C.byref(g), tr["msg_plan"].data_ptr(), sm)
        for i in range(S):
            if tc_msg:
                _lib.call("imp_edge_messages_grouped_tc32_planned", C.byref(g), tr["msg_plan"].data_ptr(), h[i].data_ptr(), d,
                          tab.data_ptr() + 4 * per * i, tab.data_ptr() + 4 * per * (S + i), 0, msg.data_ptr(), sm)
            else:
                _lib.call("imp_edge_messages_grouped", C.byref(g), h[i].data_ptr(), d, tab.data_ptr() + 4 * per * i,
                          tab.data_ptr() + 4 * per * (S + i), 0, msg.data_ptr(), cws.data_ptr(), sm)
            _lib.call("imp_segment_sum", C.byref(g), msg.data_ptr(), d, agg[i].data_ptr(), sm)
            wc, wa = self._gru_struct("cat", i), self._gru_struct("an", i)
            # tensor-core forward (3xTF32, csrc/fwd_tc32.cu) where it is built (atom_dim 32); the fp32 SIMT kernels otherwise
            tc_fwd = d == 32 and getattr(self, "tc_forward", True)
            if gates is None:
                if tc_fwd:
                    _lib.call("imp_gated_update_tc32", h[i].data_ptr(), agg[i].data_ptr(), N, batch.n_cat_atoms, d, C.byref(wc),
                              C.byref(wa), C.c_float(self.LN_EPS), h[i + 1].data_ptr(), None, None, None, sm)
                else:
                    _lib.call("imp_gated_update", h[i].data_ptr(), agg[i].data_ptr(), N, batch.n_cat_atoms, d, C.byref(wc),
                              C.byref(wa), C.c_float(self.LN_EPS), h[i + 1].data_ptr(), sm)
            else:  # the gates are kept for the backward pass (it then skips the recomputation of the three Dense layers)
                _lib.call("imp_gated_update_tc32" if tc_fwd else "imp_gated_update_train", h[i].data_ptr(), agg[i].data_ptr(), N,
                          batch.n_cat_atoms, d, C.byref(wc),
                          C.byref(wa), C.c_float(self.LN_EPS), h[i + 1].data_ptr(), gates[i][0].data_ptr(), gates[i][1].data_ptr(),
                          gates[i][2].data_ptr(), sm)
        pooled, dpooled = self._buf("tr_pooled", 2 * P * d), self._buf("tr_dpooled", 2 * P * d)
        _lib.call("imp_global_sum_pool", batch.dev["mol_ptr"].data_ptr(), batch.dev["atom_id"].data_ptr(), 2 * P,
                  h[S].data_ptr(), d, pooled.data_ptr(), sm)
        return dict(h=h, agg=agg, gates=gates, pooled=pooled, dpooled=dpooled, tab=tab, per=per, msg=msg, cws=cws, tr=tr, g=g)

    def _backward_base(self, batch, kept, occurrence_norms=True, first_step=0):
        """From kept["dpooled"] back through GlobalSumPool and the message-passing steps ``first_step .. S-1`` (and the
        Embeddings when first_step == 0) into the flat gradient bucket.  ``first_step`` > 0 serves partially frozen models
        (train_melting_point_transfer.py:206-217): nothing below the first trainable step is differentiated."""
        import torch

        s = self.spec
        d, S, K, Vb = s["atom_dim"], s["num_steps"], s["bond_dim"], s["bond_vocab_size"]
        st = self._train_state()
        h, agg, dpooled, tab, per, msg, cws, tr, g = (kept[k] for k in ("h", "agg", "dpooled", "tab", "per", "msg", "cws", "tr", "g"))
        N, P = batch.n_atoms, batch.n_pairs
        sm = _stream()
        G = st["g"]
        lib = _lib.load()
        # ---- back through the pooling and the S steps
        ga, gb = self._buf("tr_ga", N * d), self._buf("tr_gb", N * d)
        occ_bond = occurrence_norms and self.bond_occurrence_supported()
        daggs = [self._buf(f"tr_dagg{i}" if occ_bond else "tr_dagg", N * d) for i in range(S)]  # kept per step for the norm
        _lib.call("imp_pool_bwd", batch.dev["mol_ptr"].data_ptr(), batch.dev["atom_id"].data_ptr(), 2 * P, dpooled.data_ptr(), d,
                  ga.data_ptr(), sm)
        ws_gru = self._buf("tr_ws_gru", lib.imp_gated_update_bwd_workspace_floats(d))
        dtable = self._buf("tr_dtable", 2 * per)
        ws_dt = self._buf("tr_ws_dt", max(1, tr["n_chunks"]) * d * d)
        G["bond_emb"].zero_()
        cur, nxt = ga, gb
        for i in reversed(range(first_step, S)):
            dagg = daggs[i]
            wc, wa = self._gru_struct("cat", i), self._gru_struct("an", i)
            gc = G[f"cat_gu_{i}.dense_z.kernel"].data_ptr()  # the layer's 8 variables are contiguous from here
            gn = G[f"an_gu_{i}.dense_z.kernel"].data_ptr()
            if kept.get("gates") is None:
                _lib.call("imp_gated_update_bwd", h[i].data_ptr(), agg[i].data_ptr(), cur.data_ptr(), N, batch.n_cat_atoms, d,
                          C.byref(wc), C.byref(wa), C.c_float(self.LN_EPS), nxt.data_ptr(), dagg.data_ptr(), gc, gn,
                          ws_gru.data_ptr(), sm)
            else:
                zt, rt, tt = kept["gates"][i]
                _lib.call("imp_gated_update_bwd_tc" if getattr(self, "tc_backward", True) else "imp_gated_update_bwd_stored", h[i].data_ptr(), agg[i].data_ptr(), zt.data_ptr(), rt.data_ptr(), tt.data_ptr(),
                          cur.data_ptr(), N, batch.n_cat_atoms, d, C.byref(wc), C.byref(wa), C.c_float(self.LN_EPS), nxt.data_ptr(),
                          dagg.data_ptr(), gc, gn, ws_gru.data_ptr(), sm)
            # dh += sum over the (symmetric) live entries of mult * T[b]^T dagg[src]
            if d == 32 and tr.get("msg_plan") is not None and getattr(self, "tc_backward", True):
                _lib.call("imp_edge_messages_grouped_tc32_planned", C.byref(g), tr["msg_plan"].data_ptr(), dagg.data_ptr(), d,
                          tab.data_ptr() + 4 * per * i, tab.data_ptr() + 4 * per * (S + i), 1, msg.data_ptr(), sm)
            else:
                _lib.call("imp_edge_messages_grouped", C.byref(g), dagg.data_ptr(), d, tab.data_ptr() + 4 * per * i,
                          tab.data_ptr() + 4 * per * (S + i), 1, msg.data_ptr(), cws.data_ptr(), sm)
            _lib.call("imp_segment_sum_add", C.byref(g), msg.data_ptr(), d, nxt.data_ptr(), sm)
            _lib.call("imp_bond_transform_bwd", C.byref(g), tr["entry_dst"].data_ptr(), tr["chunk_begin"].data_ptr(),
                      tr["chunk_end"].data_ptr(), tr["n_chunks"], tr["bucket_chunk_ptr"].data_ptr(), dagg.data_ptr(),
                      h[i].data_ptr(), d, K, self._ptr("bond_emb"), self._ptr(f"cat_bmm_{i}.bond_transform"),
                      self._ptr(f"an_bmm_{i}.bond_transform"), G[f"cat_bmm_{i}.bond_transform"].data_ptr(),
                      G[f"an_bmm_{i}.bond_transform"].data_ptr(), G["bond_emb"].data_ptr(), dtable.data_ptr(), ws_dt.data_ptr(), sm)
            cur, nxt = nxt, cur
        if first_step > 0:
            return
        ws_e = self._buf("tr_ws_emb", lib.imp_embed_bwd_workspace_floats(s["atom_vocab_size"], d))
        _lib.call("imp_embed_bwd", batch.dev["atom_id"].data_ptr(), cur.data_ptr(), N, s["atom_vocab_size"], d,
                  G["atom_emb"].data_ptr(), ws_e.data_ptr(), sm)
        # ---- per-occurrence squared norms of the two Embedding gradients (csrc/occ_norm.cu) -> bucket tail [2], [3]
        tail = st["tail"]
        if occurrence_norms:
            ws_n = self._buf("tr_ws_norm", max(1024, lib.imp_bond_occurrence_norm2_workspace_floats(batch.n_unique)))
            _lib.call("imp_sumsq", cur.data_ptr(), N * d, tail.data_ptr() + 8, ws_n.data_ptr(), sm)
            if occ_bond:
                ob = lib.imp_occ_pack_bytes(d, K)
                opk = self._buf("tr_occ_packed", ob * 2 * S, torch.uint8)
                for ti, t in enumerate(TOWERS):
                    for i in range(S):
                        _lib.call("imp_occ_pack", self._ptr(f"{t}_bmm_{i}.bond_transform"), d, K, opk.data_ptr() + ob * (ti * S + i), sm)
                hp = (C.c_void_p * S)(*[h[i].data_ptr() for i in range(S)])
                dp = (C.c_void_p * S)(*[daggs[i].data_ptr() for i in range(S)])
                _lib.call("imp_bond_occurrence_norm2", C.byref(g), tr["n_cat_unique"], tr["entry_dst"].data_ptr(), S, hp, dp, d, K,
                          opk.data_ptr(), opk.data_ptr() + ob * S, tail.data_ptr() + 12, ws_n.data_ptr(), sm)

    def loss_and_grads(self, batch, global_batch=None, occurrence_norms=True):
        """Forward + backward on a packed batch with ``batch.target``.  Returns (sse, out): device tensors holding this
        rank's sum of squared errors and predictions; gradients of mean-squared-error over ``global_batch`` pairs
        (default: this batch; ``"sum"``: un-normalised sum-gradients, for a count that is all-reduced with them) are left
        in the flat bucket ``self._train['grad']`` (WITHOUT the l2 terms, which the optimizer kernel adds after the
        all-reduce so that they are counted once).  With ``occurrence_norms`` the tail of the bucket receives the
        per-occurrence squared norms of the two Embedding gradients (same scaling as the gradients)."""
        import torch

        s = self.spec
        d, S = s["atom_dim"], s["num_steps"]
        fp, mix = s["fp_size"], s["mixing_size"]
        visc = s["kind"] == "viscosity"
        fp2 = 0 if visc else fp
        st = self._train_state()
        kept = self._forward_kept(batch)
        P = batch.n_pairs
        sm = _stream()
        G = st["g"]
        lib = _lib.load()
        pooled, dpooled = kept["pooled"], kept["dpooled"]
        # ---- loss + readout backward
        out = torch.empty(P, dtype=torch.float32, device=self.device)
        rc, ra = self._readout_struct("cat"), self._readout_struct("an")
        nh = fp2 if fp2 else 3
        n_ro = 2 * (d * fp + fp + fp * mix + mix) + mix * nh + nh + ((fp2 + 1) if fp2 else 0)
        ro = self._buf("tr_ro_grads", n_ro)
        ws_ro = self._buf("tr_ws_ro", lib.imp_readout_bwd_workspace_floats(d, fp, mix, fp2))
        scale = 2.0 if global_batch == "sum" else 2.0 / float(global_batch if global_batch else P)
        if visc:
            if batch.dev_T is None:
                raise ValueError("viscosity model needs batch.temperature")
            W1, b1, W2, b2, Tp = self._ptr("head.kernel"), self._ptr("head.bias"), None, None, batch.dev_T.data_ptr()
        else:
            W1, b1, W2, b2, Tp = (self._ptr("head1.kernel"), self._ptr("head1.bias"), self._ptr("head2.kernel"),
                                  self._ptr("head2.bias"), None)
        _lib.call("imp_readout_bwd", pooled.data_ptr(), P, d, fp, mix, fp2, C.byref(rc), C.byref(ra), W1, b1, W2, b2, Tp,
                  batch.dev_y.data_ptr(), C.c_float(scale), dpooled.data_ptr(), out.data_ptr(), ro.data_ptr(),
                  st["sse"].data_ptr(), ws_ro.data_ptr(), sm)
        o = 0
        for t in TOWERS:
            for name, cnt in ((f"{t}_fp.kernel", d * fp), (f"{t}_fp.bias", fp), (f"{t}_mix.kernel", fp * mix),
                              (f"{t}_mix.bias", mix)):
                G[name].view(-1).copy_(ro[o:o + cnt])
                o += cnt
        heads = (("head.kernel", mix * 3), ("head.bias", 3)) if visc else (
            ("head1.kernel", mix * fp2), ("head1.bias", fp2), ("head2.kernel", fp2), ("head2.bias", 1))
        for name, cnt in heads:
            G[name].view(-1).copy_(ro[o:o + cnt])
            o += cnt
        self._backward_base(batch, kept, occurrence_norms=occurrence_norms)
        return st["sse"], out

    def gradients(self):
        """Host copy of the gradient bucket as {variable: array} (after loss_and_grads)."""
        return {k: v.detach().cpu().numpy().copy() for k, v in self._train_state()["g"].items()}

    def train_step(self, batch, lr=1e-3, clipnorm=1.0, beta1=0.9, beta2=0.999, eps=1e-7, global_batch=None, group=None,
                   embedding_clip="occurrence"):
        """One optimiser step (train_viscosity.py:227-230).  With an initialised process group the gradient bucket -- whose
        tail carries the squared-error sum, this rank's pair count and the per-occurrence Embedding norms -- is
        all-reduced ONCE; ranks may hold different numbers of pairs (the mean is taken over the all-reduced count).
        ``global_batch`` fixes the divisor instead (e.g. gradient accumulation over several calls).  Returns the loss
        (mse + l2, weights before the update, as Keras reports it) as a device scalar tensor -- no host synchronisation."""
        import torch.distributed as dist

        if embedding_clip not in ("occurrence", "dense"):
            raise ValueError("embedding_clip must be 'occurrence' or 'dense'")
        st = self._train_state()
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        occ = embedding_clip == "occurrence"
        summed = world > 1 and global_batch is None
        self.loss_and_grads(batch, global_batch="sum" if summed else (global_batch or batch.n_pairs), occurrence_norms=occ)
        tail = st["tail"]
        if summed:
            tail[1:2].fill_(float(batch.n_pairs))
        if world > 1:
            dist.all_reduce(st["grad"], op=dist.ReduceOp.SUM, group=group)
        st["step"] += 1
        n_occ = 0 if not occ else (2 if self.bond_occurrence_supported() else 1)
        _lib.call("imp_clip_adam_sparse", self.flat.data_ptr(), st["grad"].data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(),
                  st["var_off"].data_ptr(), st["var_l2"].data_ptr(), len(self.var_names), st["norms2"].data_ptr(),
                  C.c_float(clipnorm if clipnorm else 0.0), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                  st["step"], n_occ, st["occ_var"].data_ptr(), tail.data_ptr() + 8, tail.data_ptr() + 4 if summed else None,
                  tail.data_ptr(), C.c_float(1.0 / float(global_batch or batch.n_pairs)), st["loss"].data_ptr(), _stream())
        self._tables_valid = False
        return st["loss"][0].clone()

    # ------------------------------------------------------------------------------------------ fit / evaluate
    def evaluate(self, records, label=None):
        """Keras ``model.evaluate``: mean squared error over ``records`` (src/dataset.py schema, labels under
        ``label``) plus the l2 regularisation terms, with the current weights (fp32 kernels)."""
        from .graph import pack_records

        label = label or ("log_eta" if self.spec["kind"] == "viscosity" else "mp")
        batch = pack_records(records, self.spec["bond_vocab_size"], label=label)
        pred = self.forward_packed(batch.to(self.device))
        st = self._train_state() if self.spec["atom_dim"] == 32 else None
        nv = len(self.var_names)
        if st is None:  # shapes without backward kernels still evaluate: the same kernel, its own small state
            import torch

            l2 = l2_terms(self.spec)
            offs = [self.var_off[k] for k in self.var_names] + [self.flat.numel()]
            st = {"var_off": torch.tensor(offs, dtype=torch.int64, device=self.device),
                  "var_l2": torch.tensor([l2.get(k, 0.0) for k in self.var_names], dtype=torch.float32, device=self.device)}
        scratch = self._buf("eval_scratch", nv + 2)
        _lib.call("imp_eval_loss", pred.data_ptr(), batch.dev_y.data_ptr(), batch.n_pairs, self.flat.data_ptr(),
                  st["var_off"].data_ptr(), st["var_l2"].data_ptr(), nv, scratch.data_ptr(), scratch.data_ptr() + 4 * (nv + 1), _stream())
        return float(scratch[nv + 1].item())

    def fit(self, records, validation_data=None, epochs=1, batch_size=32, shuffle=True, patience=None,
            restore_best_weights=True, label=None, seed=0, verbose=0, lr=1e-3, clipnorm=1.0):
        """The reference's training loop (train_viscosity.py:328-338: ``model.fit(x, y, validation_data, epochs,
        batch_size=32, callbacks=[EarlyStopping(monitor="val_loss", patience=50, restore_best_weights=True)])``) on
        record lists: per epoch a fresh shuffle, mini-batches of ``batch_size`` (the last one may be smaller), one
        ``train_step`` each; epoch loss = sample-weighted mean of the batch losses [Keras semantics]; ``val_loss`` =
        ``evaluate(validation_data)`` at the end of the epoch; early stopping on ``val_loss`` with ``patience`` (None =
        off), strict improvement, optional restore of the best epoch's weights.  Returns ``{"loss": [...],
        "val_loss": [...]}`` like ``History.history``.  Mini-batches are packed on the host (imp_pack_host)."""
        from .graph import pack_records

        label = label or ("log_eta" if self.spec["kind"] == "viscosity" else "mp")
        rng = np.random.default_rng(seed)
        n = len(records)
        hist = {"loss": [], "val_loss": []} if validation_data is not None else {"loss": []}
        best, best_w, wait = float("inf"), None, 0
        for epoch in range(epochs):
            order = rng.permutation(n) if shuffle else np.arange(n)
            tot, losses = 0, []
            for lo in range(0, n, batch_size):
                idx = order[lo:lo + batch_size]
                batch = pack_records([records[i] for i in idx], self.spec["bond_vocab_size"], label=label)
                losses.append((len(idx), self.train_step(batch.to(self.device), lr=lr, clipnorm=clipnorm)))
                tot += len(idx)
            hist["loss"].append(float(sum(k * float(l.item()) for k, l in losses) / max(tot, 1)))
            if validation_data is not None:
                v = self.evaluate(validation_data, label)
                hist["val_loss"].append(v)
                if v < best:
                    best, wait = v, 0
                    if restore_best_weights:
                        best_w = self.flat.clone()
                else:
                    wait += 1
                if verbose:
                    print(f"Epoch {epoch + 1}/{epochs} - loss: {hist['loss'][-1]:.6f} - val_loss: {v:.6f}")
                if patience is not None and wait >= patience:
                    break
            elif verbose:
                print(f"Epoch {epoch + 1}/{epochs} - loss: {hist['loss'][-1]:.6f}")
        if best_w is not None and restore_best_weights and validation_data is not None and hist["val_loss"][-1] > best:
            self.flat.copy_(best_w)
            self._tables_valid = False
        return hist
