"""TEST INFRASTRUCTURE ONLY -- torch (CPU) restatement of the reference's MPNN forward on the
reference's own PADDED inputs, op for op, with every intermediate exposed.

Follows (paths relative to /root/reference):
  * BondMatrixMessage.call   models/layers.py:100-117  (gather, tensordot to a (B,E,d,d) temporary,
                                                        batched mat-vec, src>0 & tgt>0 mask)
  * Reduce.call              models/layers.py:57-83    (boolean_mask + scatter_nd onto (batch, tgt), tgt>0)
  * GatedUpdate.call         models/layers.py:142-156  (z, r, h~, blend, LayerNorm eps=1e-3, residual)
  * GlobalSumPool.call       models/layers.py:161-164  (masked sum, atom_id > 0)
  * SliceParamA/B/C, ScaleTemperature, ComputeLogEta   models/layers.py:10-42
  * viscosity graph          train_viscosity.py:139-231
  * melting-point graph      train_melting_point.py:137-215 (bond_dim = atom_dim**2, 2-layer head)
  * loss / optimizer         train_viscosity.py:227-230 (mse + l2 on the fingerprint kernels,
                             Adam(1e-3, clipnorm=1.0) -> per-variable clip, [Keras semantics])

dtype is a parameter: float64 is the checker, float32 mirrors what TF computes and is the
``cpu_baseline`` that bench.py times (kind "port": TensorFlow itself is not installable here).
Pinned against tests/golden/*.npz (reference source executed under oracle/tf_shim.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

TOWERS = ("cat", "an")


def make_spec(kind="viscosity", atom_vocab_size=124, bond_vocab_size=72, atom_dim=32, bond_dim=8,
              fp_size=32, mixing_size=20, num_steps=4):
    """Hyper-parameters with the defaults of build_model (train_viscosity.py:139-147;
    train_melting_point.py:137-146 where bond_dim is forced to atom_dim**2)."""
    if kind == "melting_point":
        bond_dim = atom_dim * atom_dim
    return dict(kind=kind, atom_vocab_size=atom_vocab_size, bond_vocab_size=bond_vocab_size,
                atom_dim=atom_dim, bond_dim=bond_dim, fp_size=fp_size, mixing_size=mixing_size,
                num_steps=num_steps)


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


def l2_terms(spec):
    """(parameter name, coefficient) of the kernel regularisers (train_viscosity.py:189;
    train_melting_point.py:172,196)."""
    if spec["kind"] == "viscosity":
        return [("cat_fp.kernel", 1e-4), ("an_fp.kernel", 1e-4)]
    return [("cat_fp.kernel", 1e-5), ("an_fp.kernel", 1e-5), ("head1.kernel", 1e-5)]


def init_params(spec, seed=1, trained_like=False, bond_scale=1.0):
    """Keras default initialisers ([Keras semantics]) drawn from numpy default_rng(seed):
    Embedding U(-0.05,0.05); glorot_uniform (rank-3: fans multiplied by prod(shape[:-2])); zero
    biases; LayerNorm gamma 1 / beta 0.  ``trained_like`` perturbs biases / gamma / beta so that
    they are exercised; ``bond_scale`` multiplies bond_transform (sensitivity)."""
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
            if leaf == "bond_transform":
                w = w * bond_scale
        elif leaf == "gamma":
            w = np.ones(shp) + (rng.normal(0, 0.1, size=shp) if trained_like else 0)
        else:  # bias / beta
            w = rng.normal(0, 0.1, size=shp) if trained_like else np.zeros(shp)
        out[name] = np.ascontiguousarray(w, dtype=np.float64)
    return out


def to_torch(params, dtype=torch.float64, requires_grad=False):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in params.items()}


# ------------------------------------------------------------------------------ layers
def bond_matrix_message(atom_state, bond_state, conn, W):
    """models/layers.py:100-117."""
    src_idx, tgt_idx = conn[:, :, 0], conn[:, :, 1]
    src_atoms = torch.gather(atom_state, 1, src_idx[:, :, None].expand(-1, -1, atom_state.shape[2]))
    bond_mats = torch.tensordot(bond_state, W, dims=([2], [0]))  # (B,E,d,d), the 4 KiB/edge temporary
    messages = torch.matmul(bond_mats, src_atoms[..., None])[..., 0]
    valid = (src_idx > 0) & (tgt_idx > 0)
    return messages * valid[..., None].to(messages.dtype)


def reduce_messages(messages, tgt_idx, atom_ref):
    """models/layers.py:57-83 -- scatter-add onto (batch, tgt) for tgt > 0."""
    B, E, d = messages.shape
    N = atom_ref.shape[1]
    flat_idx = (torch.arange(B)[:, None] * N + tgt_idx).reshape(-1)
    valid = tgt_idx.reshape(-1) > 0
    out = torch.zeros(B * N, d, dtype=messages.dtype)
    out = out.index_add(0, flat_idx[valid], messages.reshape(-1, d)[valid])
    return out.reshape(B, N, d)


def gated_update(atom_state, agg, p, prefix, eps=1e-3):
    """models/layers.py:142-156 (dropout rate 0 -> identity)."""
    concat = torch.cat([atom_state, agg], dim=-1)
    z = torch.sigmoid(concat @ p[f"{prefix}.dense_z.kernel"] + p[f"{prefix}.dense_z.bias"])
    r = torch.sigmoid(concat @ p[f"{prefix}.dense_r.kernel"] + p[f"{prefix}.dense_r.bias"])
    h_input = torch.cat([r * atom_state, agg], dim=-1)
    h_tilde = torch.tanh(h_input @ p[f"{prefix}.dense_h.kernel"] + p[f"{prefix}.dense_h.bias"])
    new_state = (1 - z) * atom_state + z * h_tilde
    mean = new_state.mean(dim=-1, keepdim=True)
    var = ((new_state - mean) ** 2).mean(dim=-1, keepdim=True)  # biased, like tf.nn.moments
    normed = (new_state - mean) * torch.rsqrt(var + eps)
    normed = normed * p[f"{prefix}.layernorm.gamma"] + p[f"{prefix}.layernorm.beta"]
    return normed + atom_state


def global_sum_pool(atom_features, atom_ids):
    """models/layers.py:161-164."""
    mask = (atom_ids > 0).to(atom_features.dtype)[..., None]
    return (atom_features * mask).sum(dim=1)


def forward(spec, p, x, keep=False, taps=None):
    """Whole graph on the padded input dict ``x`` (numpy int32 arrays, train_viscosity.py:306-314).
    ``p``: dict of torch tensors.  Returns (out (B,1), intermediates dict).  ``taps``: optional dict that receives
    the two Embedding outputs of every tower (``{t}_atom_rows`` (B,N,d), ``{t}_bond_rows`` (B,E,K)) with
    ``retain_grad`` set: their gradients are the per-occurrence rows of the IndexedSlices gradient TensorFlow
    hands to the optimizer for the two Embedding variables (see ``adam_step``)."""
    inter = {}
    dt = p["atom_emb"].dtype
    S = spec["num_steps"]
    fps = {}
    for t in TOWERS:
        atom_ids = torch.as_tensor(np.asarray(x[f"{t}_atom"]), dtype=torch.long)
        bond_ids = torch.as_tensor(np.asarray(x[f"{t}_bond"]), dtype=torch.long)
        conn = torch.as_tensor(np.asarray(x[f"{t}_connectivity"]), dtype=torch.long)
        h = p["atom_emb"][atom_ids]
        b = p["bond_emb"][bond_ids]
        if taps is not None:
            if h.requires_grad:
                h.retain_grad(), b.retain_grad()
            taps[f"{t}_atom_rows"], taps[f"{t}_bond_rows"] = h, b
        if keep:
            inter[f"{t}_h_0"] = h
        for i in range(S):
            m = bond_matrix_message(h, b, conn, p[f"{t}_bmm_{i}.bond_transform"])
            agg = reduce_messages(m, conn[:, :, 1], h)
            h = gated_update(h, agg, p, f"{t}_gu_{i}")
            if keep:
                inter[f"{t}_agg_{i}"] = agg
                inter[f"{t}_h_{i + 1}"] = h
        pool = global_sum_pool(h, atom_ids)
        fp = torch.relu(pool @ p[f"{t}_fp.kernel"] + p[f"{t}_fp.bias"])
        fps[t] = fp
        if keep:
            inter[f"{t}_pool"] = pool
            inter[f"{t}_fp"] = fp
    mixed = sum(torch.relu(fps[t] @ p[f"{t}_mix.kernel"] + p[f"{t}_mix.bias"]) for t in TOWERS)
    if keep:
        inter["mixed"] = mixed
    if spec["kind"] == "viscosity":
        vp = mixed @ p["head.kernel"] + p["head.bias"]
        A = vp[:, 0:1]
        Bp = torch.clamp(torch.nn.functional.softplus(vp[:, 1:2]), 0.0, 20.0)
        C = torch.clamp(torch.nn.functional.softplus(vp[:, 2:3]), 0.1, 50.0)
        T = torch.as_tensor(np.asarray(x["temperature"]), dtype=torch.float32).to(dt) / 100.0
        out = A + Bp / (T + C + 1e-6)
        if keep:
            inter["visc_params"] = vp
    else:
        hid = torch.relu(mixed @ p["head1.kernel"] + p["head1.bias"])
        out = hid @ p["head2.kernel"] + p["head2.bias"]
    return out, inter


def predict(spec, params, x, dtype=torch.float64, batch_size=None, keep=False):
    """numpy in / numpy out convenience.  ``batch_size`` mimics Keras ``predict`` slicing (default 32
    in the reference, train_viscosity.py:366); None = one batch."""
    p = to_torch(params, dtype)
    n = len(x["cat_atom"])
    bs = n if batch_size is None else batch_size
    outs, inters = [], []
    with torch.no_grad():
        for s in range(0, n, bs):
            xb = {k: v[s : s + bs] for k, v in x.items()}
            o, it = forward(spec, p, xb, keep=keep)
            outs.append(o.numpy())
            inters.append({k: v.numpy() for k, v in it.items()})
    out = np.concatenate(outs, 0)
    if keep:
        return out, {k: np.concatenate([it[k] for it in inters], 0) for k in inters[0]}
    return out


def loss_fn(spec, p, x, y, taps=None):
    """Keras compiled loss: mean((y - yhat)^2) with y (B,) expanded to (B,1), plus the l2 kernel
    regularisers (train_viscosity.py:189,227-230)."""
    out, _ = forward(spec, p, x, taps=taps)
    yt = torch.as_tensor(np.asarray(y), dtype=out.dtype).reshape(-1, 1)
    loss = ((yt - out) ** 2).mean()
    for name, coef in l2_terms(spec):
        loss = loss + coef * (p[name] ** 2).sum()
    return loss, out


def loss_and_grads(spec, params, x, y, dtype=torch.float64, occurrence_norms=False):
    """Loss, dense gradients of every variable and predictions.  With ``occurrence_norms`` a fourth value is
    returned: {"atom_emb": s_a, "bond_emb": s_b}, the sum of squares over the UN-deduplicated per-occurrence
    gradient rows of the two Embedding variables (every (tower, sample, atom slot) / (tower, sample, edge slot) of
    the padded inputs is one occurrence) -- what ``tf.clip_by_norm`` sees for an IndexedSlices gradient."""
    p = to_torch(params, dtype, requires_grad=True)
    taps = {} if occurrence_norms else None
    loss, out = loss_fn(spec, p, x, y, taps=taps)
    loss.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
    if not occurrence_norms:
        return float(loss.detach()), grads, out.detach().numpy()
    occ = {"atom_emb": float(sum((taps[f"{t}_atom_rows"].grad ** 2).sum() for t in TOWERS)),
           "bond_emb": float(sum((taps[f"{t}_bond_rows"].grad ** 2).sum() for t in TOWERS))}
    return float(loss.detach()), grads, out.detach().numpy(), occ


def adam_step(params, grads, m, v, step, lr=1e-3, clipnorm=1.0, beta1=0.9, beta2=0.999, eps=1e-7,
              occurrence_norm2=None):
    """[Keras semantics] ``Adam(1e-3, clipnorm=1.0)`` of TF / Keras 2.12 (environment.yml:10; the library source is
    not under /root/reference, so this is restated from keras/optimizers/optimizer.py and adam.py of that release):

      _BaseOptimizer.apply_gradients:  grads = self._clip_gradients(grads)           # per variable:
                                                                                      #   tf.clip_by_norm(g, clipnorm)
                                       grads = self._deduplicate_sparse_grad(grads)  # AFTER the clip
      tf.clip_by_norm(t, c):           values = t.values if IndexedSlices else t
                                       l2norm = sqrt(sum(values * values));  values * c / max(l2norm, c)
      Adam.update_step:                alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
                                       m += (g - m) (1 - b1);  v += (g^2 - v) (1 - b2)
                                       w -= alpha * m / (sqrt(v) + eps)
                                       (sparse branch: the same decay of ALL rows of m and v, the de-duplicated
                                        rows scattered in, and a dense update of w -- identical to the dense branch)

    The only place where the two ``Embedding`` variables (train_viscosity.py:163-164; gradient = IndexedSlices whose
    rows are the per-occurrence gradients, concatenated over the two towers) differ from dense variables is therefore
    the CLIP NORM: it is taken over the un-deduplicated rows, sqrt(sum_occurrences |g_occ|^2), not over the summed
    dense gradient.  ``occurrence_norm2`` = {variable: that sum of squares} (from ``loss_and_grads(...,
    occurrence_norms=True)``) selects it; variables not listed (and ``None``) use the dense norm.
    ``step`` is 1-based.  Updates in place, returns the per-variable norms used for clipping."""
    norms = {}
    alpha = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    for k in params:
        g = np.asarray(grads[k], dtype=np.float64)
        nrm = float(np.sqrt((g * g).sum()))
        if occurrence_norm2 is not None and k in occurrence_norm2:
            nrm = float(np.sqrt(occurrence_norm2[k]))
        norms[k] = nrm
        if clipnorm is not None:
            g = g * (clipnorm / max(nrm, clipnorm))
        m[k] += (g - m[k]) * (1 - beta1)
        v[k] += (g * g - v[k]) * (1 - beta2)
        params[k] -= alpha * m[k] / (np.sqrt(v[k]) + eps)
    return norms
