"""CPU: the packed-CSR specification (oracle/ref_pack.py) against the reference's padded arrays."""
import numpy as np
import pytest

from conftest import GOLDEN_F64, load_golden
from ionic_mpnn_b200 import synth
from oracle import ref_inputs, ref_pack


@pytest.mark.parametrize("name", GOLDEN_F64)
def test_packed_equals_padded_live_set(name):
    meta, x, _, _, _ = load_golden(name)
    pk = ref_pack.pack_records(meta["records"], meta["spec"]["bond_vocab_size"])
    for tower, t in enumerate(("cat", "an")):
        assert ref_pack.packed_live_multiset(pk, tower) == ref_pack.padded_live_multiset(x, t)
    live = sum((x[f"{t}_connectivity"] > 0).all(-1).sum() for t in ("cat", "an"))
    assert pk["n_edges"] == live


def test_quirks_multiplicity_two_and_atom_zero_isolated():
    recs = synth.make_records(20, seed=5)
    pk = ref_pack.pack_records(recs, 72)
    assert set((pk["edge_bm"] >> 16).tolist()) == {2}
    starts = pk["mol_ptr"][:-1]
    deg = np.diff(pk["row_ptr"])
    assert (deg[starts] == 0).all()  # first atom of every ion receives nothing
    assert not np.isin(pk["col_src"], starts).any()  # ... and sends nothing
    assert pk["n_atoms"] == sum(r["cation"]["num_atoms"] + r["anion"]["num_atoms"] for r in recs)


def test_rows_sorted_and_buckets_consistent():
    recs = synth.make_records(30, seed=6, skewed=True)
    pk = ref_pack.pack_records(recs, 72)
    rp, cs, bm = pk["row_ptr"], pk["col_src"], pk["edge_bm"] & 0xFFFF
    for v in range(pk["n_atoms"]):
        keys = list(zip(bm[rp[v]:rp[v + 1]].tolist(), cs[rp[v]:rp[v + 1]].tolist()))
        assert keys == sorted(keys) and len(set(keys)) == len(keys)
    bp, perm = pk["bucket_ptr"], pk["bucket_perm"]
    assert bp[0] == 0 and bp[-1] == pk["n_unique"] and sorted(perm.tolist()) == list(range(pk["n_unique"]))
    dst = np.repeat(np.arange(pk["n_atoms"]), np.diff(rp))
    for g in range(2 * 72):
        e = perm[bp[g]:bp[g + 1]]
        assert (np.diff(e) > 0).all()
        assert ((dst[e] >= pk["n_cat_atoms"]) == (g >= 72)).all() and (bm[e] == g % 72).all()


def test_truncation_matches_reference_padding():
    recs = synth.make_records(6, seed=7, n_min=4, n_max=10)
    max_edges = 6  # shorter than most ions' edge lists -> "e[:max_len]" truncation kicks in
    conn = {}
    x = {}
    for t, key in (("cat", "cation"), ("an", "anion")):
        e, b = ref_inputs.preprocess_edges_and_bonds([r[key]["edge_indices"] for r in recs],
                                                     [[v + 1 for v in r[key]["bond_ids"]] for r in recs], max_edges)
        x[f"{t}_connectivity"], x[f"{t}_bond"] = e, b
    pk = ref_pack.pack_records(recs, 72, max_edges=max_edges)
    for tower, t in enumerate(("cat", "an")):
        assert ref_pack.packed_live_multiset(pk, tower) == ref_pack.padded_live_multiset(x, t)


def test_empty_and_edgeless():
    recs = [{"cation": {"atom_ids": [3], "bond_ids": [], "edge_indices": [], "num_atoms": 1},
             "anion": {"atom_ids": [1, 2], "bond_ids": [4, 4], "edge_indices": [(0, 1), (1, 0)], "num_atoms": 2}, "T": 300.0}]
    pk = ref_pack.pack_records(recs, 72)
    assert pk["n_unique"] == 0 and pk["n_edges"] == 0 and pk["n_atoms"] == 3  # the only bond touches atom 0
    pk0 = ref_pack.pack_records([], 72)
    assert pk0["n_atoms"] == 0 and pk0["row_ptr"].tolist() == [0] and pk0["mol_ptr"].tolist() == [0]
    with pytest.raises(ValueError):
        ref_pack.pack_records([{"cation": {"atom_ids": [0, 1], "bond_ids": [0, 0], "edge_indices": [(1, 2), (2, 1)]},
                                "anion": {"atom_ids": [0], "bond_ids": [], "edge_indices": []}}], 72)
