"""CPU: the standalone oracle (oracle/ref_model.py, oracle/ref_inputs.py) against the vectors produced
by executing the reference's own source (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN_F64, load_golden
from oracle import ref_inputs, ref_model


@pytest.mark.parametrize("name", GOLDEN_F64)
def test_weights_reproducible(name):
    meta, _, _, _, params = load_golden(name)
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(np.ascontiguousarray(params[k], dtype=np.float64).tobytes())
    assert h.hexdigest() == meta["weights_sha256_f64"]
    assert {k: tuple(v.shape) for k, v in params.items()} == ref_model.param_shapes(meta["spec"])


@pytest.mark.parametrize("name", GOLDEN_F64)
def test_padded_inputs_bit_exact(name):
    meta, x, _, _, _ = load_golden(name)
    mine = ref_inputs.build_inputs(meta["records"], with_temperature=meta["kind"] == "viscosity")
    assert set(mine) == set(x)
    for k in x:
        assert mine[k].dtype == x[k].dtype and mine[k].shape == x[k].shape, k
        assert np.array_equal(mine[k], x[k]), k


@pytest.mark.parametrize("name", GOLDEN_F64)
def test_forward_fp64_matches_reference_source(name):
    meta, x, inter, out, params = load_golden(name)
    mine, mi = ref_model.predict(meta["spec"], params, x, dtype=torch.float64, keep=True)
    np.testing.assert_allclose(mine, out, rtol=1e-11, atol=1e-11)
    for k, v in inter.items():
        if "_msg_" in k:
            continue
        np.testing.assert_allclose(mi[k], v, rtol=1e-10, atol=1e-11, err_msg=k)


def test_messages_match_reference_source():
    meta, x, inter, _, params = load_golden("visc_small")
    p = ref_model.to_torch(params)
    for t in ref_model.TOWERS:
        h = p["atom_emb"][torch.as_tensor(x[f"{t}_atom"], dtype=torch.long)]
        b = p["bond_emb"][torch.as_tensor(x[f"{t}_bond"], dtype=torch.long)]
        conn = torch.as_tensor(x[f"{t}_connectivity"], dtype=torch.long)
        m = ref_model.bond_matrix_message(h, b, conn, p[f"{t}_bmm_0.bond_transform"])
        np.testing.assert_allclose(m.numpy(), inter[f"{t}_msg_0"], rtol=1e-11, atol=1e-13)


def test_batched_predict_is_batch_invariant():
    meta, x, _, out, params = load_golden("visc_batched_predict")
    a = ref_model.predict(meta["spec"], params, x, batch_size=32)
    b = ref_model.predict(meta["spec"], params, x, batch_size=None)
    np.testing.assert_allclose(a, out, rtol=1e-11)
    np.testing.assert_allclose(b, out, rtol=1e-11)


def test_fp32_reference_wiring_close_to_fp64():
    _, x, _, out32, _ = load_golden("visc_default_init_f32")
    meta, _, _, out64, params = load_golden("visc_default_init")
    np.testing.assert_allclose(out32, out64, rtol=2e-5, atol=2e-5)
    mine32 = ref_model.predict(meta["spec"], params, x, dtype=torch.float32)
    np.testing.assert_allclose(mine32, out64, rtol=2e-5, atol=2e-5)


# ---- known-answer tests authored from the code's semantics ----
def _tiny_spec(**kw):
    base = dict(atom_vocab_size=6, bond_vocab_size=4, atom_dim=2, bond_dim=2, fp_size=2, mixing_size=2, num_steps=1)
    base.update(kw)
    return ref_model.make_spec("viscosity", **base)


def test_three_atom_chain_by_hand():
    """Chain 0-1-2, bond ids (shifted) 1 and 2.  Atom 0 neither sends nor receives; the 1<->2 bond is
    seen twice in each direction, so agg[1] = 2*A_2 h_2 and agg[2] = 2*A_2 h_1."""
    spec = _tiny_spec()
    rec = [{"cation": {"atom_ids": [0, 1, 2], "bond_ids": [0, 0, 1, 1], "edge_indices": [(0, 1), (1, 0), (1, 2), (2, 1)],
                       "num_atoms": 3},
            "anion": {"atom_ids": [3], "bond_ids": [], "edge_indices": [], "num_atoms": 1}, "T": 300.0}]
    x = ref_inputs.build_inputs(rec)
    assert x["cat_connectivity"].shape == (1, 8, 2)
    params = ref_model.init_params(spec, seed=0)
    params["atom_emb"] = np.array([[9, 9], [1, 0], [0, 1], [1, 1], [2, 3], [5, 7]], float)
    params["bond_emb"] = np.array([[9, 9], [1, 0], [0, 1], [1, 1]], float)
    W = np.array([[[1, 2], [3, 4]], [[5, 6], [7, 8]]], float)  # W[k,l,m]
    params["cat_bmm_0.bond_transform"] = W
    _, inter = ref_model.predict(spec, params, x, keep=True)
    h = params["atom_emb"][[1, 2, 3]]  # shifted ids 1,2,3
    A2 = W[1]  # shifted bond id 2 -> bond_emb row [0,1] -> W[1]
    want = np.stack([np.zeros(2), 2 * A2 @ h[2], 2 * A2 @ h[1]])
    np.testing.assert_allclose(inter["cat_agg_0"][0, :3], want, rtol=0, atol=1e-12)
    np.testing.assert_allclose(want, [[0, 0], [22, 30], [12, 16]])


def test_padding_is_invisible():
    recs = [r for r in load_golden("visc_small")[0]["records"]]
    meta, _, _, out, params = load_golden("visc_small")
    x_big = ref_inputs.build_inputs(recs, max_atoms=40, max_edges=90)
    got = ref_model.predict(meta["spec"], params, x_big)
    np.testing.assert_allclose(got, out, rtol=1e-12)


def test_layernorm_epsilon_matters():
    meta, x, inter, _, params = load_golden("visc_default_init")
    p = ref_model.to_torch(params)
    h0 = p["atom_emb"][torch.as_tensor(x["cat_atom"], dtype=torch.long)]
    agg = torch.as_tensor(inter["cat_agg_0"])
    a = ref_model.gated_update(h0, agg, p, "cat_gu_0", eps=1e-3)
    b = ref_model.gated_update(h0, agg, p, "cat_gu_0", eps=1e-5)
    np.testing.assert_allclose(a.numpy(), inter["cat_h_1"], rtol=1e-10, atol=1e-12)
    assert float((a - b).abs().max()) > 0.1  # row variance ~8e-4 < eps: eps visibly matters


def test_head_clip_boundaries():
    spec = _tiny_spec()
    params = ref_model.init_params(spec, seed=0)
    rec = [{"cation": {"atom_ids": [0], "bond_ids": [], "edge_indices": [], "num_atoms": 1},
            "anion": {"atom_ids": [1], "bond_ids": [], "edge_indices": [], "num_atoms": 1}, "T": 250.0}]
    x = ref_inputs.build_inputs(rec)
    params["head.kernel"] = np.zeros((2, 3))
    for b1, b2, Bw, Cw in [(50.0, 100.0, 20.0, 50.0), (-50.0, -50.0, 0.0, 0.1), (1.0, 1.0, np.log1p(np.e), np.log1p(np.e))]:
        params["head.bias"] = np.array([0.5, b1, b2])
        got = ref_model.predict(spec, params, x)[0, 0]
        np.testing.assert_allclose(got, 0.5 + Bw / (2.5 + Cw + 1e-6), rtol=1e-9)


def test_gradients_finite_difference():
    meta, x, _, _, params = load_golden("visc_small")
    spec = meta["spec"]
    y = np.array([r["log_eta"] for r in meta["records"]])
    loss, grads, _ = ref_model.loss_and_grads(spec, params, x, y)
    rng = np.random.default_rng(0)
    for name in ["atom_emb", "bond_emb", "cat_bmm_1.bond_transform", "an_gu_0.dense_h.kernel",
                 "cat_gu_2.layernorm.gamma", "cat_fp.kernel", "an_mix.bias", "head.kernel"]:
        w = params[name]
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in w.shape)
            eps = 1e-6
            old = w[idx]
            w[idx] = old + eps
            lp = ref_model.loss_and_grads(spec, params, x, y)[0]
            w[idx] = old - eps
            lm = ref_model.loss_and_grads(spec, params, x, y)[0]
            w[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - grads[name][idx]) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, grads[name][idx])
