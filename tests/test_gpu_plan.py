"""GPU (B200): the tile plan (imp_fused_plan) and the planned fused forward (imp_mpnn_forward_fused_planned, generation 5)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL16 = 2e-2  # north-star tolerance of the 16-bit tensor path


def _rel(got, want):
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    return float((np.abs(got - want) / np.maximum(np.abs(want), 1.0)).max())


def _plan(batch, compact=False, atom_vocab=124):
    from ionic_mpnn_b200 import _lib

    lib = _lib.load()
    nb = lib.imp_fused_plan_bytes(batch.n_pairs, batch.n_atoms, batch.n_unique, batch.max_mol_atoms)
    assert nb >= 256
    plan = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    if compact:
        cb = batch.to_compact("cuda")
        cg = cb.compact_struct()
        _lib.call("imp_fused_plan", None, C.byref(cg), atom_vocab, batch.max_mol_atoms, plan.data_ptr(), nb, None)
    else:
        batch.to("cuda")
        g = batch.c_struct()
        _lib.call("imp_fused_plan", C.byref(g), None, atom_vocab, batch.max_mol_atoms, plan.data_ptr(), nb, None)
    torch.cuda.synchronize()
    return plan.cpu().numpy()


@pytest.mark.parametrize("n_pairs,seed,lo,hi", [(1, 1, 10, 40), (300, 2, 10, 40), (1000, 3, 10, 40), (257, 4, 1, 12),
                                                (100, 5, 60, 128), (513, 6, 1, 128)])
def test_plan_records_are_a_faithful_cut_of_the_csr_batch(n_pairs, seed, lo, hi):
    from ionic_mpnn_b200 import graph
    from oracle import ref_plan

    batch, _, _ = graph.synth_batch(n_pairs, seed=seed, n_min=lo, n_max=hi)
    stats = None
    for compact in (False, True):
        raw = _plan(batch, compact)
        st = ref_plan.check_plan(raw, batch.host, batch.n_pairs, 124, 72)
        if stats is not None:  # both input feeds describe the same graph: same tiles (their order inside the plan may differ)
            assert st == stats
        stats = st
    if (lo, hi) == (10, 40) and n_pairs >= 300:
        assert stats["fill"] >= 0.93, stats  # best-fit over 256-molecule windows (the contiguous cut reaches ~0.88)
    print(n_pairs, lo, hi, stats)


def test_plan_degenerate_molecules_and_multiplicities():
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.graph import FlatIons
    from oracle import ref_plan

    ions = [{"atom_ids": [5], "bond_ids": [], "edge_indices": [], "num_atoms": 1},
            {"atom_ids": [1, 2, 3, 4], "bond_ids": [], "edge_indices": [], "num_atoms": 4},
            {"atom_ids": [7, 8, 9], "bond_ids": [3, 3, 3, 3, 4, 4], "num_atoms": 3,
             "edge_indices": [(1, 2), (2, 1), (1, 2), (2, 1), (0, 1), (1, 0)]}] * 30
    f = FlatIons.from_ion_dicts(ions)
    batch = graph.pack_flat(f, f, 72)
    raw = _plan(batch)
    st = ref_plan.check_plan(raw, batch.host, batch.n_pairs, 124, 72)
    assert st["tiles"] >= 2  # at most 32 molecules per tile


def test_plan_refuses_batches_outside_its_envelope():
    """> 31 entries in a row: the status word says so, and MPNNModel routes such a batch to the self-contained kernel."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model
    from oracle import ref_plan

    n = 40
    star = {"atom_ids": list(range(1, n + 1)), "num_atoms": n, "bond_ids": [], "edge_indices": []}
    for i in range(2, n):  # atom 1 is bonded to every other atom but atom 0
        star["edge_indices"] += [(1, i), (i, 1)]
        star["bond_ids"] += [i % 70, i % 70]
    recs = [{"pair_id": k, "cation": star, "anion": star, "T": 300.0, "log_eta": 1.0} for k in range(5)]
    batch = graph.pack_records(recs, 72)
    assert batch.max_in_degree > 31
    hdr, _ = ref_plan.parse(_plan(batch))
    assert hdr["status"] == 1
    ref = build_model(124, 72, precision="fp32", seed=3)
    fz = build_model(124, 72, precision="fp16", seed=3, fused=True)
    assert not fz.planned_supported(batch)
    want = ref.predict(batch)
    got = fz.predict(batch)
    assert _rel(got, want) <= RTOL16


@pytest.mark.parametrize("n_pairs,seed,lo,hi", [(1, 1, 10, 40), (64, 2, 10, 40), (3000, 3, 10, 40), (300, 4, 1, 128)])
def test_planned_forward_matches_the_self_contained_kernel_and_the_fp32_path(n_pairs, seed, lo, hi):
    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.viscosity import build_model

    batch, _, _ = graph.synth_batch(n_pairs, seed=seed, n_min=lo, n_max=hi)
    batch.to("cuda")
    ref = build_model(124, 72, precision="fp32", seed=3)
    want32 = ref.forward_packed(batch).cpu().numpy()
    for precision, pflags, gen in (("fp16", 0, 6), ("fp16_precise", 0, 6), ("fp16", _lib.TC_GEN5, 5), ("fp16", 0, 7),
                                   ("fp16_precise", 0, 7), ("fp16", 0, 8), ("fp16_precise", 0, 8)):  # kernel generations 6, 5, 7, 8
        planned = build_model(124, 72, precision=precision, seed=3, fused=True)
        planned.extra_tc_flags = pflags
        planned.fused_gen = gen
        assert planned.planned_gen() == gen
        gen3 = build_model(124, 72, precision=precision, seed=3, fused=True)
        gen3.use_plan = False
        assert planned.planned_supported(batch) and not gen3.planned_supported(batch)
        a = planned.forward_packed(batch).cpu().numpy()
        b = gen3.forward_packed(batch).cpu().numpy()
        c = planned.forward_packed(batch.to_compact("cuda")).cpu().numpy()
        planned.check_status()
        assert np.array_equal(a, c), "the compact feed and the int32 CSR feed give the same plan contents"
        if batch.narrow_ok:
            c16 = planned.forward_packed(batch.to_compact("cuda", narrow=True)).cpu().numpy()
            assert np.array_equal(a, c16), "so does the narrow compact feed (16-bit entry words)"
        assert np.array_equal(a, planned.forward_packed(batch).cpu().numpy()), "run-to-run bit-identical"
        # the same arithmetic up to the LayerNorm evaluation order; an fp32 rounding difference can flip the 16-bit rounding
        # of an operand, so the two kernels agree to a few operand ulps (2^-11), not to fp32 rounding
        assert _rel(a, b) <= 5e-3, _rel(a, b)
        assert _rel(a, want32) <= RTOL16
    assert planned.launches_per_forward(batch) == 4
    assert _lib.load().imp_fused_plan_bytes(-1, 0, 0, 0) < 0


def test_plan_capacity_overflow_is_reported_and_predict_retries():
    """plan_slack sizes the plan buffer for well-filled tiles; a batch that needs more tiles reports status 2, the model
    falls back to the safe bound and predict() repeats the call."""
    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.viscosity import build_model

    batch, _, _ = graph.synth_batch(600, seed=8, n_min=65, n_max=70)  # one molecule per tile: 2x the tiles plan_slack expects
    ref = build_model(124, 72, precision="fp32", seed=3)
    want = ref.predict(batch)
    fz = build_model(124, 72, precision="fp16", seed=3, fused=True)
    fz.plan_slack = 1.25  # an explicit factor is taken as it is ("auto", the default, would see max_mol_atoms = 70 and allow 2.2x)
    fz.forward_packed(batch)
    torch.cuda.synchronize()
    with pytest.raises(_lib.PlanCapacityError):
        fz.check_status()
    assert fz.plan_slack is None
    fz2 = build_model(124, 72, precision="fp16", seed=3, fused=True)
    fz2.plan_slack = 1.25
    got = fz2.predict(batch)  # retries by itself
    assert _rel(got, want) <= RTOL16
    fz3 = build_model(124, 72, precision="fp16", seed=3, fused=True)  # default: sized from the largest molecule, no overflow
    got3 = fz3.forward_packed(batch).cpu().numpy()
    fz3.check_status()
    assert np.array_equal(got3.reshape(-1), got.reshape(-1))


def test_predict_stream_picks_the_narrow_feed_and_matches_predict():
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    chunks = [graph.synth_batch(700, seed=40 + i)[0] for i in range(3)]
    m = build_model(124, 72, precision="fp16", seed=3, fused=True)
    out, nbytes = m.predict_stream(chunks)
    torch.cuda.synchronize()
    m.check_status()
    assert m._stream_state["fields"][3] == "edge_h"
    want = np.concatenate([m.predict(c).reshape(-1) for c in chunks])
    assert np.array_equal(out.numpy(), want)
    per_pair = nbytes / sum(c.n_pairs for c in chunks)
    assert per_pair < 340, per_pair  # 16-bit atom words + 16-bit entry words + offsets + temperature


def test_large_atom_vocabulary_reads_the_embedding_table_from_global_memory():
    """The planned forward stages the atom-embedding table in shared memory when it fits (<= ~160 rows); a larger vocabulary
    takes the global-memory read path of the same kernel."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    cat = graph.synth_flat(400, 5, 10, 40, atom_types=599, bond_types=71)
    an = graph.synth_flat(400, 6, 10, 40, atom_types=599, bond_types=71)
    T = np.random.default_rng(3).uniform(273.15, 373.15, 400).astype(np.float32)
    b = graph.pack_flat(cat, an, 72, temperature=T).to("cuda")
    ref = build_model(600, 72, precision="fp32", seed=4)
    fz = build_model(600, 72, precision="fp16", seed=4, fused=True)
    assert fz.planned_supported(b)
    want = ref.forward_packed(b).cpu().numpy()
    got = fz.forward_packed(b).cpu().numpy()
    fz.check_status()
    assert _rel(got, want) <= RTOL16
