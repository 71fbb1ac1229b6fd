"""GPU (B200): imp_pack_device against the host packer (itself bit-exact against oracle/ref_pack.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CSR = ("mol_ptr", "atom_id", "row_ptr", "col_src", "edge_bm")


def _compare(cat, an, **kw):
    from ionic_mpnn_b200 import graph

    host = graph.pack_flat(cat, an, 72, **{k: v for k, v in kw.items() if k != "compact"})
    dev = graph.pack_flat_device(cat, an, 72, **kw)
    assert (dev.n_pairs, dev.n_atoms, dev.n_cat_atoms, dev.n_unique, dev.n_edges) == \
           (host.n_pairs, host.n_atoms, host.n_cat_atoms, host.n_unique, host.n_edges)
    for k in CSR:
        got = dev.dev[k].cpu().numpy()[: len(host.host[k])]
        assert np.array_equal(got, host.host[k]), k
    if kw.get("compact", True):
        host.compact()
        assert np.array_equal(dev.dev["mol_eptr"].cpu().numpy(), host.chost["mol_eptr"])
        assert np.array_equal(dev.dev["atom_w"].cpu().numpy().view(np.uint16)[: host.n_atoms], host.chost["atom_w"])
        assert np.array_equal(dev.dev["edge_w"].cpu().numpy().view(np.uint32)[: host.n_unique], host.chost["edge_w"])
    return host, dev


def test_device_packer_matches_host_packer_on_synthetic_ions():
    from ionic_mpnn_b200 import graph

    for seed, n, skew in ((1, 1000, False), (2, 257, True), (3, 1, False)):
        cat = graph.synth_flat(n, 2 * seed + 1, skewed=skew)
        an = graph.synth_flat(n, 2 * seed + 2, skewed=skew)
        _compare(cat, an)
    big_c, big_a = graph.synth_flat(50, 11, 40, 120), graph.synth_flat(50, 12, 40, 120)
    _compare(big_c, big_a)


def test_device_packer_truncation_undoubled_and_degenerate_inputs():
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.graph import FlatIons

    cat, an = graph.synth_flat(300, 5), graph.synth_flat(300, 6)
    _compare(cat, an, max_edges=17)          # the reference's truncation to 2 * max_edges entries
    _compare(cat, an, double_edges=False)    # inputs that already carry both directions once
    pre = [FlatIons(f.atom_ptr, f.atom_ids + 1, f.edge_ptr, f.edge_src, f.edge_dst, f.bond_ids + 1) for f in (cat, an)]
    _compare(pre[0], pre[1], shift_ids=False, compact=False)   # ids already shifted by the caller (pack_padded path)
    ions = [{"atom_ids": [5], "bond_ids": [], "edge_indices": [], "num_atoms": 1},
            {"atom_ids": [1, 2, 3, 4], "bond_ids": [], "edge_indices": [], "num_atoms": 4},
            {"atom_ids": [7, 8, 9], "bond_ids": [3, 3, 3, 3, 4, 4], "num_atoms": 3,
             "edge_indices": [(1, 2), (2, 1), (1, 2), (2, 1), (0, 1), (1, 0)]}]   # duplicates -> multiplicity 4; atom 0 masked
    f = FlatIons.from_ion_dicts(ions)
    host, dev = _compare(f, f)
    assert (dev.dev["edge_bm"].cpu().numpy()[: host.n_unique] >> 16).max() == 4


def test_device_packer_reports_bad_indices_and_oversized_molecules():
    from ionic_mpnn_b200 import _lib, graph
    from ionic_mpnn_b200.graph import FlatIons

    bad = FlatIons.from_ion_dicts([{"atom_ids": [1, 2], "bond_ids": [0], "edge_indices": [(1, 5)], "num_atoms": 2}])
    with pytest.raises(_lib.ImpError):
        graph.pack_flat_device(bad, bad, 72)
    n = 40
    dense = {"atom_ids": list(range(n)), "num_atoms": n, "bond_ids": [], "edge_indices": []}
    for a in range(n):
        for b in range(a + 1, n):
            dense["edge_indices"] += [(a, b), (b, a)]
            dense["bond_ids"] += [1, 1]
    big = FlatIons.from_ion_dicts([dense])     # 3120 doubled entries > 512: the device packer refuses, the host packs
    with pytest.raises(_lib.ImpError):
        graph.pack_flat_device(big, big, 72)
    assert graph.pack_flat(big, big, 72).n_unique > 0


def test_forward_on_device_packed_batch_is_identical():
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.viscosity import build_model

    cat, an = graph.synth_flat(2000, 31), graph.synth_flat(2000, 32)
    T = np.random.default_rng(0).uniform(273.15, 373.15, 2000).astype(np.float32)
    host = graph.pack_flat(cat, an, 72, temperature=T).to("cuda")
    dev = graph.pack_flat_device(cat, an, 72, temperature=T)
    m = build_model(124, 72, precision="fp16", seed=2)
    want = m.forward_packed(host).cpu().numpy()
    assert np.array_equal(m.forward_packed(dev).cpu().numpy(), want)
    assert np.array_equal(m.forward_packed(dev.as_compact()).cpu().numpy(), want)
    m32 = build_model(124, 72, precision="fp32", seed=2)
    assert np.array_equal(m32.forward_packed(dev).cpu().numpy(), m32.forward_packed(host).cpu().numpy())


def test_device_packed_batch_through_the_staged_tensor_path():
    """A device-packed batch has no bond buckets: the staged tensor path then runs the CSR-order message kernel instead of
    the bucketed tcgen05 GEMM (melting-point model: no fused kernel for bond_dim 1024).  Same predictions within tolerance."""
    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    cat, an = graph.synth_flat(500, 31), graph.synth_flat(500, 32)
    host = graph.pack_flat(cat, an, 72)
    dev = graph.pack_flat_device(cat, an, 72)
    m = MPNNModel(make_spec("melting_point"), seed=2, precision="fp16")
    a = m.forward_packed(host).cpu().numpy()
    b = m.forward_packed(dev).cpu().numpy()
    assert np.isfinite(b).all()
    # fp16 grouped tensor GEMM vs fp32 SIMT messages, both inside the 2e-2 class of the tensor path (north_star)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1.0)) <= 1e-2
