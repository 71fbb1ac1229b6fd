"""GPU (B200): the tcgen05 / TMEM plumbing in isolation (imp_tc_selftest) against a CPU product."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


def _tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def run_selftest(N, K, kind, swap=0, seed=0):
    from ionic_mpnn_b200 import _lib

    rng = np.random.default_rng(seed)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dD = torch.zeros(128, N, dtype=torch.float32, device="cuda")
    _lib.call("imp_tc_selftest", dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), N, K, kind, swap,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return A, B, dD.cpu().numpy()


@pytest.mark.parametrize("N,K", [(32, 32), (64, 64), (32, 256), (64, 16)])
def test_bf16_mma_matches_cpu(N, K):
    A, B, D = run_selftest(N, K, 0)
    want = _bf16_round(A).astype(np.float64) @ _bf16_round(B).astype(np.float64).T
    assert np.abs(D - want).max() <= 1e-3 * max(1.0, np.abs(want).max()), np.abs(D - want).max()


@pytest.mark.parametrize("N,K", [(32, 32), (64, 64), (32, 8)])
def test_tf32_mma_matches_cpu(N, K):
    A, B, D = run_selftest(N, K, 1)
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    # tf32 keeps 10 mantissa bits: error per product ~2^-11 relative (round or truncate, hardware's choice)
    assert np.abs(D - exact).max() <= 4e-3 * np.sqrt(K) * 4.0, np.abs(D - exact).max()
    # and an exactly representable input must give the exact answer
    A2, B2 = _tf32_trunc(A), _tf32_trunc(B)
    from ionic_mpnn_b200 import _lib

    dA, dB = torch.from_numpy(A2).cuda(), torch.from_numpy(B2).cuda()
    dD = torch.zeros(128, N, dtype=torch.float32, device="cuda")
    _lib.call("imp_tc_selftest", dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), N, K, 1, 0,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = A2.astype(np.float64) @ B2.astype(np.float64).T
    assert np.abs(dD.cpu().numpy() - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


# ---------------------------------------------------------------------------- bf16 tensor-core model path
BF16_RTOL = 2e-2  # north_star: "2e-2 relative (bf16 tensor-core path) on log_eta / mp predictions"


def _rel(got, want):
    """The north-star metric, plain: max |got - want| / max(|want|, 1) over the predictions."""
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))


@pytest.mark.parametrize("precision", ["bf16_precise", "bf16", "fp16"])
@pytest.mark.parametrize("name", ["visc_default_init", "visc_trained_like"])
def test_bf16_path_matches_reference_golden(name, precision):
    from conftest import load_golden
    from ionic_mpnn_b200.viscosity import build_model

    meta, x, inter, out, params = load_golden(name)
    model = build_model(124, 72, precision=precision, fused=False)
    model.set_weights(params)
    got = model.predict(x)
    scale = float(np.abs(out).max())
    err = _rel(got, out)
    print(f"{name} {precision}: max rel err {err:.3e}, scaled {np.abs(got - out).max() / scale:.3e}")
    if name == "visc_default_init":
        assert err <= BF16_RTOL, err
    else:  # sensitivity weights (bond_transform x10): tolerance relative to the prediction scale
        assert np.abs(got - out).max() / scale <= BF16_RTOL


def test_bf16_gated_update_vs_fp32_kernel_per_step():
    """Same inputs through imp_gated_update (fp32 SIMT) and imp_gated_update_tc (tcgen05): h after each step."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    batch, _, _ = graph.synth_batch(3000, seed=5)
    a = build_model(124, 72, precision="fp32", seed=3)
    b = build_model(124, 72, precision="bf16_precise", seed=3)
    c = build_model(124, 72, precision="fp16", seed=3)
    batch.to("cuda")
    _, ia = a.forward_packed(batch, keep=True)
    _, ib = b.forward_packed(batch, keep=True)
    _, ic = c.forward_packed(batch, keep=True)
    torch.cuda.synchronize()
    for i in range(1, 5):
        ref = ia["h"][i]
        scale = float(ref.abs().max())
        eb = float((ib["h"][i] - ref).abs().max()) / scale
        ec = float((ic["h"][i] - ref).abs().max()) / scale
        print(f"step {i}: |h| max {scale:.3f}  bf16_precise err {eb:.3e}  bf16 err {ec:.3e}")
        assert eb <= 2e-2 and ec <= 2e-2
    # ragged tail: a tower whose atom count is not a multiple of 128 is handled (rows masked, no OOB write)
    assert batch.n_cat_atoms % 128 != 0 or (batch.n_atoms - batch.n_cat_atoms) % 128 != 0


@pytest.mark.parametrize("kind,precision", [("viscosity", "fp16"), ("viscosity", "bf16"), ("melting_point", "fp16")])
def test_tensor_message_kernel_vs_fp32_kernel(kind, precision):
    """imp_edge_messages_tc (csrc/msg_tc.cu: bond-type-grouped tcgen05 GEMM over gathered source rows) + imp_segment_sum
    against imp_message_agg (fp32 SIMT) on the SAME atom states, for every step's table of both towers."""
    import ctypes as C

    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    spec = make_spec(kind)
    batch, _, _ = graph.synth_batch(1800, seed=21, with_temperature=(kind == "viscosity"))
    batch.to("cuda")
    ref = MPNNModel(spec, seed=3, precision="fp32")
    _, ia = ref.forward_packed(batch, keep=True)
    m = MPNNModel(spec, seed=3, precision=precision, fused=False)
    m.refresh_tables()
    g = batch.c_struct()
    S, d = spec["num_steps"], 32
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert batch.n_atoms > 148 * 256 and batch.n_cat_atoms % 256 != 0
    for i in range(S):
        h = ia["h"][i].contiguous()
        agg = torch.full_like(h, float("nan"))
        msg = torch.full((batch.n_unique, d), float("nan"), device="cuda")
        cws = torch.empty(2 * 72 + 1, dtype=torch.int32, device="cuda")
        base = m._ws["msg_packed"].data_ptr()
        _lib.call("imp_edge_messages_tc", C.byref(g), h.data_ptr(), d, base + m._msg_pack_bytes * i,
                  base + m._msg_pack_bytes * (S + i), m.tc_flags(), msg.data_ptr(), cws.data_ptr(), st)
        _lib.call("imp_segment_sum", C.byref(g), msg.data_ptr(), d, agg.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.isfinite(msg).all(), "every CSR entry must have been written exactly once"
        want = ia["agg"][i]
        err = float((agg - want).abs().max() / want.abs().max())
        print(f"{kind} {precision} step {i}: |agg| max {float(want.abs().max()):.3f}, tensor message kernel err {err:.3e}")
        assert torch.isfinite(agg).all()
        assert err <= (2e-3 if precision == "fp16" else 1.6e-2), err
        # rows without live entries (the first atom of every ion) are exactly zero
        deg = torch.from_numpy(np.diff(batch.host["row_ptr"])).cuda()
        assert float(agg[deg == 0].abs().max()) == 0.0


def test_reduce_folded_into_gated_update_is_bit_identical():
    """imp_reduce_gated_update_tc (message rows summed in the load stage) == imp_segment_sum + imp_gated_update_tc."""
    import ctypes as C

    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    spec = make_spec("melting_point")
    batch, _, _ = graph.synth_batch(1500, seed=23, with_temperature=False)
    batch.to("cuda")
    m = MPNNModel(spec, seed=3, precision="fp16", fused=False)
    m.refresh_tables()
    g = batch.c_struct()
    d, S = 32, spec["num_steps"]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = torch.Generator(device="cuda").manual_seed(1)
    h = torch.randn(batch.n_atoms, d, device="cuda", generator=rng)
    msg = torch.randn(batch.n_unique, d, device="cuda", generator=rng)
    agg = torch.empty_like(h)
    out_a, out_b = torch.empty_like(h), torch.empty_like(h)
    gb = m._ws["gru_packed"].data_ptr()
    _lib.call("imp_segment_sum", C.byref(g), msg.data_ptr(), d, agg.data_ptr(), st)
    _lib.call("imp_gated_update_tc", h.data_ptr(), agg.data_ptr(), batch.n_atoms, batch.n_cat_atoms, d, gb,
              gb + m._gru_pack_bytes * S, C.c_float(1e-3), m.tc_flags(), out_a.data_ptr(), st)
    _lib.call("imp_reduce_gated_update_tc", C.byref(g), h.data_ptr(), msg.data_ptr(), d, gb, gb + m._gru_pack_bytes * S,
              C.c_float(1e-3), m.tc_flags(), out_b.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.isfinite(out_b).all()
    assert torch.equal(out_a, out_b)
    # and the model's routes agree: keep=True (separate kernels, fp32 messages) == folded with fp32 messages -- the atom states bit
    # for bit (above); the predictions to fp32 rounding, because keep=True reads out with the double-accumulating
    # imp_pool_head_* (it returns the intermediates) and the plain forward with imp_global_sum_pool + the fp32 imp_readout_*;
    # the default folded route (16-bit message rows and gathers) within the tensor path's tolerance
    a, _ = m.forward_packed(batch, keep=True)
    m.fp32_messages = True
    b = m.forward_packed(batch)
    assert torch.allclose(a, b, rtol=2e-6, atol=2e-6 * float(a.abs().max()))
    m.fp32_messages = False
    c = m.forward_packed(batch)
    err = float(((c - a).abs() / a.abs().clamp(min=1.0)).max())
    print(f"16-bit message rows vs fp32 message rows: max rel err {err:.3e}")
    assert torch.isfinite(c).all() and err <= 5e-3


def test_bf16_cfg1_thousand_pairs_vs_fp64_oracle():
    from ionic_mpnn_b200 import synth
    from ionic_mpnn_b200.viscosity import build_model
    from oracle import ref_inputs, ref_model

    recs = synth.make_records(1000, seed=0)
    spec = ref_model.make_spec("viscosity")
    params = ref_model.init_params(spec, seed=1)
    want = ref_model.predict(spec, params, ref_inputs.build_inputs(recs), batch_size=32)
    # staged tensor kernels (fused=False).  IEEE-half operands meet the north-star tolerance; bfloat16 operands
    # (8-bit significand, the same tcgen05 rate) are measured and bounded more loosely: the weight rounding is a
    # systematic error that the 25-atom pooled sums amplify.
    for precision, tol in (("fp16", BF16_RTOL), ("bf16", 1.5e-1)):
        model = build_model(124, 72, precision=precision, fused=False)
        model.set_weights(params)
        got = model.predict(recs)
        err = _rel(got, want)
        print(f"{precision} staged path, 1000 pairs: max rel err {err:.3e}")
        assert err <= tol, (precision, err)


# ---------------------------------------------------------------------------- A operand from tensor memory
@pytest.mark.parametrize("kind", [2, 3])
@pytest.mark.parametrize("N,K", [(32, 32), (64, 64), (32, 256)])
def test_ts_mma_matches_cpu(N, K, kind):
    """tcgen05.mma with A in TMEM (written by tcgen05.st, 16-bit pairs): the form the fused forward uses."""
    A, B, D = run_selftest(N, K, kind)
    rnd = _bf16_round if kind == 2 else (lambda x: x.astype(np.float16).astype(np.float32))
    want = rnd(A).astype(np.float64) @ rnd(B).astype(np.float64).T
    assert np.abs(D - want).max() <= 1e-3 * max(1.0, np.abs(want).max()), np.abs(D - want).max()


# ---------------------------------------------------------------------------- fused whole-tower forward
def _fused_vs_staged(n_pairs, seed, precision, skewed=False, n_min=10, n_max=40, tc_flags=0):
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    batch, _, _ = graph.synth_batch(n_pairs, seed=seed, n_min=n_min, n_max=n_max, skewed=skewed)
    batch.to("cuda")
    ref = build_model(124, 72, precision="fp32", seed=3)
    fz = build_model(124, 72, precision=precision, seed=3, fused=True)
    fz.extra_tc_flags = tc_flags
    want = ref.forward_packed(batch).cpu().numpy()
    got = fz.forward_packed(batch).cpu().numpy()
    torch.cuda.synchronize()
    assert int(fz._ws["status"].item()) == 0
    return got, want


@pytest.mark.parametrize("precision", ["fp16", "fp16_precise", "bf16"])
@pytest.mark.parametrize("n_pairs", [1, 7, 64, 65, 1000])
def test_fused_forward_vs_fp32_kernels(n_pairs, precision):
    got, want = _fused_vs_staged(n_pairs, 11, precision)
    err = _rel(got, want)
    print(f"fused {precision}, {n_pairs} pairs: max rel err {err:.3e}")
    assert np.isfinite(got).all()
    assert err <= (BF16_RTOL if precision.startswith("fp16") else 1.5e-1), err


@pytest.mark.parametrize("tc_flags", [8, 8 | 4, 16, 32, 512, 1024])
def test_fused_forward_kernel_generations(tc_flags):
    """Earlier generations stay covered (the default, flags 0, is the planned fifth generation): fp32 Z accumulation
    (IMP_TC_F32_ZBUILD, optionally IMP_TC_MP8), two threads per row (IMP_TC_TWO_THREADS_PER_ROW), the self-contained third
    generation (IMP_TC_GEN3), its three-context variant, and the fourth generation (IMP_TC_GEN4)."""
    got, want = _fused_vs_staged(700, 4, "fp16", tc_flags=tc_flags)
    assert _rel(got, want) <= BF16_RTOL


def test_fused_forward_large_molecules_and_skew():
    """Molecules up to 120 atoms (one or two per tile), Zipf-skewed bond types."""
    got, want = _fused_vs_staged(300, 5, "fp16", skewed=True, n_min=40, n_max=120)
    assert _rel(got, want) <= BF16_RTOL


def test_fused_forward_is_deterministic_and_graph_shape_agnostic():
    got1, _ = _fused_vs_staged(3000, 2, "fp16")
    got2, _ = _fused_vs_staged(3000, 2, "fp16")
    assert np.array_equal(got1, got2)


def test_fused_matches_reference_golden_and_oracle():
    from conftest import load_golden
    from ionic_mpnn_b200 import synth
    from ionic_mpnn_b200.viscosity import build_model
    from oracle import ref_inputs, ref_model

    meta, x, inter, out, params = load_golden("visc_default_init")
    model = build_model(124, 72, precision="fp16", fused=True)
    model.set_weights(params)
    assert _rel(model.predict(x), out) <= BF16_RTOL
    recs = synth.make_records(1000, seed=0)
    spec = ref_model.make_spec("viscosity")
    params = ref_model.init_params(spec, seed=1)
    want = ref_model.predict(spec, params, ref_inputs.build_inputs(recs), batch_size=32)
    model.set_weights(params)
    got = model.predict(recs)
    err = _rel(got, want)
    print(f"fused fp16, cfg1 1000 pairs vs fp64 oracle: max rel err {err:.3e}")
    assert err <= BF16_RTOL, err


def test_fused_refuses_shapes_outside_its_envelope():
    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.viscosity import build_model

    with pytest.raises(_lib.ImpError):
        m = build_model(124, 72, atom_dim=16, precision="fp16", fused=True)
        b, _, _ = graph.synth_batch(4, seed=1)
        m.forward_packed(b.to("cuda"))
    big, _, _ = graph.synth_batch(4, seed=1, n_min=130, n_max=140)
    m = build_model(124, 72, precision="fp16", fused=True)
    with pytest.raises(_lib.ImpError):
        m.forward_packed(big.to("cuda"))
    # 'auto' falls back to the staged tensor kernels for the same batch
    m2 = build_model(124, 72, precision="fp16")
    assert np.isfinite(m2.forward_packed(big).cpu().numpy()).all()


def test_predict_stream_matches_per_chunk_predict():
    """Double-buffered streamed prediction == chunk-by-chunk forward (bit-identical: same kernels, same order)."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    m = build_model(124, 72, precision="fp16", seed=4)
    chunks = [graph.synth_batch(n, seed=40 + i)[0] for i, n in enumerate([700, 64, 1500, 1, 333])]
    out, nbytes = m.predict_stream(chunks)
    torch.cuda.synchronize()
    want = np.concatenate([m.forward_packed(c.to("cuda")).cpu().numpy() for c in chunks])
    assert nbytes > 0 and np.array_equal(out.numpy(), want)
    # staged kernels stream too (all CSR fields are copied)
    m2 = build_model(124, 72, precision="fp32", seed=4)
    out2, _ = m2.predict_stream(chunks)
    torch.cuda.synchronize()
    want2 = np.concatenate([m2.forward_packed(c.to("cuda")).cpu().numpy() for c in chunks])
    assert np.array_equal(out2.numpy(), want2)


def _dense_records(n_pairs, seed, n_atoms=(20, 36), p_edge=0.35):
    """Ions that are far denser than molecules (in-degree up to ~15, > 768 unique entries per 128-atom tile): drives
    the fused kernel through its unstaged-entry path and through rows whose degree exceeds the sort key range."""
    rng = np.random.default_rng(seed)

    def ion():
        n = int(rng.integers(n_atoms[0], n_atoms[1] + 1))
        ei, bi = [], []
        for a in range(n):
            for b in range(a + 1, n):
                if rng.random() < p_edge:
                    bond = int(rng.integers(0, 71))
                    ei += [(a, b), (b, a)]
                    bi += [bond, bond]
        return {"atom_ids": [int(x) for x in rng.integers(0, 123, n)], "bond_ids": bi, "edge_indices": ei, "num_atoms": n}

    return [{"cation": ion(), "anion": ion(), "T": float(rng.uniform(273.15, 373.15))} for _ in range(n_pairs)]


def test_fused_forward_dense_graphs_and_degenerate_ions():
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    recs = _dense_records(40, 1)
    # degenerate ions: a single atom, and an ion without any bond
    recs[3]["cation"] = {"atom_ids": [5], "bond_ids": [], "edge_indices": [], "num_atoms": 1}
    recs[7]["anion"] = {"atom_ids": [1, 2, 3, 4], "bond_ids": [], "edge_indices": [], "num_atoms": 4}
    batch = graph.pack_records(recs, 72).to("cuda")
    assert np.diff(batch.host["row_ptr"]).max() > 7
    ref = build_model(124, 72, precision="fp32", seed=3)
    want = ref.forward_packed(batch).cpu().numpy()
    for flags in (0, 512, 1024, 32, 16, 8):
        fz = build_model(124, 72, precision="fp16", seed=3, fused=True)
        fz.extra_tc_flags = flags
        got = fz.forward_packed(batch).cpu().numpy()
        assert np.isfinite(got).all()
        assert _rel(got, want) <= BF16_RTOL, (flags, _rel(got, want))


def test_fused_forward_empty_batch():
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    empty = graph.pack_records([], 72)
    empty.temperature = np.zeros(0, np.float32)
    m = build_model(124, 72, precision="fp16", fused=True)
    assert m.predict(empty).shape == (0, 1)


def test_compact_feed_is_bit_identical():
    """imp_mpnn_forward_fused_compact (16-bit atom words, 32-bit entry words) == imp_mpnn_forward_fused, bit for bit,
    on molecule-like, large and dense graphs; predict_stream picks it up automatically."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    m = build_model(124, 72, precision="fp16", seed=4)
    cases = [graph.synth_batch(777, seed=21)[0], graph.synth_batch(90, seed=22, n_min=40, n_max=120, skewed=True)[0],
             graph.pack_records(_dense_records(30, 2), 72)]
    for b in cases:
        want = m.forward_packed(b.to("cuda")).cpu().numpy()
        got = m.forward_packed(b.to_compact("cuda")).cpu().numpy()
        assert np.array_equal(got, want)
        assert b.nbytes_compact() < 0.5 * b.nbytes()
    chunks = [graph.synth_batch(n, seed=60 + i)[0] for i, n in enumerate([500, 65, 900])]
    out_c, bytes_c = m.predict_stream(chunks, compact=True)
    torch.cuda.synchronize()
    out_c = out_c.numpy().copy()
    out_f, bytes_f = m.predict_stream(chunks, compact=False)
    torch.cuda.synchronize()
    assert np.array_equal(out_c, out_f.numpy()) and bytes_c < 0.5 * bytes_f


@pytest.mark.parametrize("num_steps,fp_size,mixing_size", [(1, 32, 20), (2, 16, 8), (3, 24, 32)])
def test_fused_forward_other_step_counts_and_readout_sizes(num_steps, fp_size, mixing_size):
    """The fused kernel takes 1..4 steps and the fast readout any fp / mix <= 32: parity against the fp64 oracle."""
    from ionic_mpnn_b200 import synth
    from ionic_mpnn_b200.viscosity import build_model
    from oracle import ref_inputs, ref_model

    recs = synth.make_records(200, seed=8)
    spec = ref_model.make_spec("viscosity", num_steps=num_steps, fp_size=fp_size, mixing_size=mixing_size)
    params = ref_model.init_params(spec, seed=5, trained_like=True)
    want = ref_model.predict(spec, params, ref_inputs.build_inputs(recs), batch_size=64)
    model = build_model(124, 72, num_steps=num_steps, fp_size=fp_size, mixing_size=mixing_size, precision="fp16", fused=True)
    model.set_weights(params)
    got = model.predict(recs)
    assert _rel(got, want) <= BF16_RTOL, _rel(got, want)
    ref32 = build_model(124, 72, num_steps=num_steps, fp_size=fp_size, mixing_size=mixing_size, precision="fp32")
    ref32.set_weights(params)
    assert _rel(ref32.predict(recs), want) <= 1e-5


def test_fused_forward_unbalanced_towers_and_small_vocabularies():
    """Tiny cations / large anions (the CTA split between towers follows the atom counts) and non-default vocabularies."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    cat = graph.synth_flat(500, 71, 2, 4, atom_types=9, bond_types=5)
    an = graph.synth_flat(500, 72, 60, 120, atom_types=9, bond_types=5)
    T = np.random.default_rng(3).uniform(273.15, 373.15, 500).astype(np.float32)
    b = graph.pack_flat(cat, an, 6, temperature=T).to("cuda")
    ref = build_model(10, 6, precision="fp32", seed=9)
    fz = build_model(10, 6, precision="fp16", seed=9, fused=True)
    want = ref.forward_packed(b).cpu().numpy()
    got = fz.forward_packed(b).cpu().numpy()
    assert _rel(got, want) <= BF16_RTOL
    assert np.array_equal(fz.forward_packed(b.to_compact("cuda")).cpu().numpy(), got)


def test_staged_tensor_path_unbalanced_towers_small_vocabularies_and_empty_buckets():
    """The bucket-grouped message kernels (tcgen05 for fp16 / bf16, exact fp32) with tiny cations / large anions, an odd bond
    vocabulary with unused bond types (empty buckets), chunks shorter than 128 slots, and degenerate ions."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.graph import FlatIons
    from ionic_mpnn_b200.viscosity import build_model

    cat = graph.synth_flat(300, 81, 2, 4, atom_types=9, bond_types=3)
    an = graph.synth_flat(300, 82, 60, 120, atom_types=9, bond_types=3)
    T = np.random.default_rng(5).uniform(273.15, 373.15, 300).astype(np.float32)
    b = graph.pack_flat(cat, an, 7, temperature=T).to("cuda")  # bond ids 1..3 of a vocabulary of 7: buckets 0, 4, 5, 6 empty
    ref = build_model(10, 7, precision="fp32", seed=9)
    ref.simt_messages = True                                      # CSR-order kernel
    want = ref.forward_packed(b).cpu().numpy()
    grouped = build_model(10, 7, precision="fp32", seed=9)        # bucket-grouped fp32 kernel
    assert _rel(grouped.forward_packed(b).cpu().numpy(), want) <= 1e-5
    for precision in ("fp16", "bf16"):
        m = build_model(10, 7, precision=precision, seed=9, fused=False)
        got = m.forward_packed(b).cpu().numpy()
        assert np.isfinite(got).all()
        assert _rel(got, want) <= BF16_RTOL, precision
    # ions without bonds and single-atom ions: no live entry at all in one tower
    ions = [{"atom_ids": [3], "bond_ids": [], "edge_indices": [], "num_atoms": 1},
            {"atom_ids": [1, 2, 3], "bond_ids": [], "edge_indices": [], "num_atoms": 3}]
    lone = FlatIons.from_ion_dicts(ions)
    other = graph.synth_flat(2, 83, 5, 9, atom_types=9, bond_types=3)
    b2 = graph.pack_flat(lone, other, 7, temperature=T[:2]).to("cuda")
    w2 = ref.forward_packed(b2).cpu().numpy()
    for precision in ("fp32", "fp16"):
        m = build_model(10, 7, precision=precision, seed=9, fused=False)
        assert _rel(m.forward_packed(b2).cpu().numpy(), w2) <= (1e-5 if precision == "fp32" else BF16_RTOL)


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_pipelined_message_kernel_matches_one_chunk_per_cta(precision):
    """imp_edge_messages_tc16: the persistent software-pipelined kernel (default) and the one-chunk-per-CTA kernel
    (IMP_TC_MSG_ONE_CHUNK_PER_CTA) write bit-identical message rows; both match the fp32-I/O kernel within 16-bit rounding.
    More chunks than resident CTAs (8 x 148), so that every CTA pipelines several chunks."""
    import ctypes as C

    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    spec = make_spec("melting_point")
    batch, _, _ = graph.synth_batch(5000, seed=29, with_temperature=False)
    batch.to("cuda")
    assert batch.n_unique > 2 * 8 * 148 * 128
    m = MPNNModel(spec, seed=3, precision=precision, fused=False)
    m.refresh_tables()
    g = batch.c_struct()
    d, S = 32, spec["num_steps"]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = torch.randn(batch.n_atoms, d, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    h16 = h.to(torch.float16 if precision == "fp16" else torch.bfloat16).contiguous()
    cws = torch.empty(2 * 72 + 1, dtype=torch.int32, device="cuda")
    base, mb = m._ws["msg_packed"].data_ptr(), m._msg_pack_bytes
    out = []
    for extra in (0, _lib.TC_MSG_ONE_CHUNK_PER_CTA):
        msg16 = torch.full((batch.n_unique, d), float("nan"), dtype=h16.dtype, device="cuda")
        _lib.call("imp_edge_messages_tc16", C.byref(g), h16.data_ptr(), d, base + mb, base + mb * (S + 1), m.tc_flags() | extra,
                  msg16.data_ptr(), cws.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.isfinite(msg16.float()).all(), "every CSR entry must have been written"
        out.append(msg16)
    assert torch.equal(out[0].view(torch.int16), out[1].view(torch.int16))
    # the planned form (per-batch index plan, what the model's forward runs)
    plan = torch.empty(_lib.load().imp_edge_messages_tc16_plan_bytes(batch.n_unique, 72), dtype=torch.uint8, device="cuda")
    _lib.call("imp_edge_messages_tc16_plan", C.byref(g), plan.data_ptr(), st)
    msg16 = torch.full((batch.n_unique, d), float("nan"), dtype=h16.dtype, device="cuda")
    _lib.call("imp_edge_messages_tc16_planned", C.byref(g), plan.data_ptr(), h16.data_ptr(), d, base + mb, base + mb * (S + 1),
              m.tc_flags(), msg16.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(out[0].view(torch.int16), msg16.view(torch.int16))
    msg32 = torch.empty(batch.n_unique, d, device="cuda")
    _lib.call("imp_edge_messages_tc", C.byref(g), h16.float().contiguous().data_ptr(), d, base + mb, base + mb * (S + 1), m.tc_flags(),
              msg32.data_ptr(), cws.data_ptr(), st)
    torch.cuda.synchronize()
    err = float((out[0].float() - msg32).abs().max() / msg32.abs().max())
    assert err <= (1e-3 if precision == "fp16" else 8e-3), err


def test_fused_full_size_properties_2m_pairs():
    """BASELINE configs[2] at bench size (2,097,152 pairs per GPU) through size-independent properties: the fused forward is
    deterministic run to run, every prediction is finite, a pair's prediction does not depend on which other pairs share the
    batch (a sub-batch of every 64th pair reproduces those rows; measured bit-identical, asserted within 16-bit rounding
    because the pooling order inside a tile follows the tile's degree sort), and the compact feed is bit-identical."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    P = 2_097_152
    model = build_model(124, 72, precision="fp16", fused=True)
    batch, cat, an = graph.synth_batch(P, seed=1003)
    batch.to("cuda")
    full = model.forward_packed(batch)
    again = model.forward_packed(batch).clone()
    torch.cuda.synchronize()
    assert full.shape == (P,) and bool(torch.isfinite(full).all())
    assert torch.equal(full, again)
    assert torch.equal(model.forward_packed(batch.to_compact("cuda")), full)
    idx = np.arange(0, P, 64)

    def take(ions, idx):
        ap, ep = ions.atom_ptr.astype(np.int64), ions.edge_ptr.astype(np.int64)
        na, ne = (ap[idx + 1] - ap[idx]), (ep[idx + 1] - ep[idx])
        new_ap = np.zeros(len(idx) + 1, np.int32)
        new_ap[1:] = np.cumsum(na)
        new_ep = np.zeros(len(idx) + 1, np.int32)
        new_ep[1:] = np.cumsum(ne)
        aidx = np.concatenate([np.arange(ap[i], ap[i + 1]) for i in idx])
        eidx = np.concatenate([np.arange(ep[i], ep[i + 1]) for i in idx])
        return graph.FlatIons(new_ap, np.ascontiguousarray(ions.atom_ids[aidx]), new_ep, np.ascontiguousarray(ions.edge_src[eidx]),
                              np.ascontiguousarray(ions.edge_dst[eidx]), np.ascontiguousarray(ions.bond_ids[eidx]))

    sub = graph.pack_flat(take(cat, idx), take(an, idx), 72, temperature=batch.temperature[idx])
    part = model.forward_packed(sub.to("cuda")).cpu().numpy()
    want = full.cpu().numpy()[idx]
    err = float(np.max(np.abs(part - want) / np.maximum(np.abs(want), 1.0)))
    print(f"2M-pair fused forward: sub-batch of every 64th pair within {err:.2e}")
    assert err <= 2e-3
