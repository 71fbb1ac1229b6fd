"""CPU: the C++ host packer (imp_pack_host through ctypes) bit-exact against oracle/ref_pack.py, and the
C-ABI library exporting every symbol include/imp_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_F64, ROOT, load_golden
from ionic_mpnn_b200 import _lib, graph, synth
from oracle import ref_inputs, ref_pack

FIELDS = graph.GRAPH_FIELDS


def assert_same(pk, ref):
    assert (pk.n_pairs, pk.n_atoms, pk.n_cat_atoms, pk.n_unique, pk.n_edges) == (
        ref["n_pairs"], ref["n_atoms"], ref["n_cat_atoms"], ref["n_unique"], ref["n_edges"])
    for k in FIELDS:
        assert pk.host[k].dtype == np.int32
        assert np.array_equal(pk.host[k], ref[k]), k


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "imp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(imp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in imp_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().imp_version() == 100


@pytest.mark.parametrize("seed,skewed", [(0, False), (1, True), (2, False)])
def test_records_bit_exact(seed, skewed):
    recs = synth.make_records(64, seed=seed, skewed=skewed)
    assert_same(graph.pack_records(recs, 72), ref_pack.pack_records(recs, 72))


def test_truncation_bit_exact():
    recs = synth.make_records(16, seed=9, n_min=4, n_max=10)
    assert_same(graph.pack_records(recs, 72, max_edges=6), ref_pack.pack_records(recs, 72, max_edges=6))


@pytest.mark.parametrize("name", GOLDEN_F64)
def test_padded_dict_gives_same_batch_as_records(name):
    meta, x, _, _, _ = load_golden(name)
    vb = meta["spec"]["bond_vocab_size"]
    assert_same(graph.pack_padded(x, vb), ref_pack.pack_records(meta["records"], vb))


def test_multithreaded_equals_single_thread():
    recs = synth.make_records(3000, seed=4, n_min=3, n_max=8)
    cat = graph.FlatIons.from_ion_dicts([r["cation"] for r in recs])
    an = graph.FlatIons.from_ion_dicts([r["anion"] for r in recs])
    a = graph.pack_flat(cat, an, 72, n_threads=1)
    b = graph.pack_flat(cat, an, 72, n_threads=7)
    for k in FIELDS:
        assert np.array_equal(a.host[k], b.host[k]), k
    assert a.n_edges == b.n_edges


def test_edge_cases_and_errors():
    empty = graph.pack_records([], 72)
    assert empty.n_atoms == 0 and empty.host["row_ptr"].tolist() == [0]
    recs = [{"cation": {"atom_ids": [3], "bond_ids": [], "edge_indices": []},
             "anion": {"atom_ids": [1, 2], "bond_ids": [4, 4], "edge_indices": [(0, 1), (1, 0)]}, "T": 300.0}]
    assert_same(graph.pack_records(recs, 72), ref_pack.pack_records(recs, 72))
    bad = [{"cation": {"atom_ids": [0, 1], "bond_ids": [0, 0], "edge_indices": [(1, 2), (2, 1)]},
            "anion": {"atom_ids": [0], "bond_ids": [], "edge_indices": []}}]
    with pytest.raises(_lib.ImpError, match="out of range"):
        graph.pack_records(bad, 72)
    bad[0]["cation"]["edge_indices"] = [(1, 1), (1, 1)]
    bad[0]["cation"]["bond_ids"] = [71, 71]  # shifted to 72 == vocabulary size
    with pytest.raises(_lib.ImpError):
        graph.pack_records(bad, 72)


def test_synth_ions_follow_the_recipe():
    ions = graph.synth_flat(2000, seed=11)
    n = np.diff(ions.atom_ptr)
    assert n.min() >= 10 and n.max() <= 40 and abs(n.mean() - 25) < 1.0
    assert ions.atom_ids.min() >= 0 and ions.atom_ids.max() == 122
    assert ions.bond_ids.min() == 0 and ions.bond_ids.max() == 70
    # featurize convention: consecutive (a,b),(b,a) pairs sharing the bond id
    assert np.array_equal(ions.edge_src[0::2], ions.edge_dst[1::2]) and np.array_equal(ions.edge_dst[0::2], ions.edge_src[1::2])
    assert np.array_equal(ions.bond_ids[0::2], ions.bond_ids[1::2])
    for i in range(50):  # tree + <=2 ring bonds, degree cap 4, local indices in range
        e0, e1 = ions.edge_ptr[i], ions.edge_ptr[i + 1]
        nb = (e1 - e0) // 2
        assert n[i] - 1 <= nb <= n[i] + 1
        deg = np.bincount(ions.edge_src[e0:e1], minlength=n[i])
        assert deg.max() <= 4 and ions.edge_src[e0:e1].max() < n[i]
    again = graph.synth_flat(2000, seed=11)
    assert np.array_equal(again.edge_src, ions.edge_src) and np.array_equal(again.atom_ids, ions.atom_ids)
    sk = graph.synth_flat(2000, seed=11, skewed=True)
    assert np.bincount(sk.bond_ids, minlength=71)[0] > 5 * np.bincount(sk.bond_ids, minlength=71)[40]


def test_compact_feed_round_trips_the_csr():
    """PackedGraphBatch.compact(): the 16/32-bit words decode back to atom_id / row_ptr / col_src / edge_bm exactly."""
    from ionic_mpnn_b200 import graph

    b, _, _ = graph.synth_batch(300, seed=9)
    b.compact()
    h, c = b.host, b.chost
    mp = h["mol_ptr"].astype(np.int64)
    deg = (c["atom_w"] >> 8).astype(np.int64)
    assert np.array_equal(c["atom_w"] & 0xFF, h["atom_id"])
    row_ptr = np.concatenate([[0], np.cumsum(deg)])
    assert np.array_equal(row_ptr, h["row_ptr"])
    assert np.array_equal(c["mol_eptr"], h["row_ptr"][mp])
    base = np.repeat(mp[:-1], np.diff(mp))[np.repeat(np.arange(b.n_atoms), deg)]
    assert np.array_equal((c["edge_w"] & 0xFF).astype(np.int64) + base, h["col_src"])
    assert np.array_equal(((c["edge_w"] >> 8) & 0xFF) | ((c["edge_w"] >> 16) << 16), h["edge_bm"])
    assert b.nbytes_compact() < 0.5 * b.nbytes()


def test_training_chunk_lists_partition_the_buckets():
    """train.prepare_training's host-built structures (entry_dst, chunk list of the (tower, bond) buckets) without a GPU:
    every bucket position is covered exactly once, in order, by chunks of at most DT_CHUNK entries."""
    from ionic_mpnn_b200 import graph, train

    b, _, _ = graph.synth_batch(400, seed=5)
    rp = b.host["row_ptr"].astype(np.int64)
    entry_dst = np.repeat(np.arange(b.n_atoms), np.diff(rp))
    assert len(entry_dst) == b.n_unique and (np.diff(entry_dst) >= 0).all()
    bp = b.host["bucket_ptr"].astype(np.int64)
    sizes = np.diff(bp)
    n_per = (sizes + train.DT_CHUNK - 1) // train.DT_CHUNK
    bcp = np.concatenate([[0], np.cumsum(n_per)])
    which = np.repeat(np.arange(len(sizes)), n_per)
    idx = np.arange(bcp[-1]) - bcp[which]
    begin = bp[which] + idx * train.DT_CHUNK
    end = np.minimum(bp[which + 1], begin + train.DT_CHUNK)
    covered = np.concatenate([np.arange(s, e) for s, e in zip(begin, end)]) if len(begin) else np.zeros(0, np.int64)
    assert np.array_equal(covered, np.arange(bp[-1]))
    assert (end - begin).max() <= train.DT_CHUNK and (end > begin).all()
    # the bucket key of every chunk is the bucket it was cut from
    key = (entry_dst[b.host["bucket_perm"]] >= b.n_cat_atoms) * b.bond_vocab + (b.host["edge_bm"][b.host["bucket_perm"]] & 0xFFFF)
    for c in range(0, len(begin), max(1, len(begin) // 50)):
        assert (key[begin[c]:end[c]] == which[c]).all()
    assert train.l2_terms({"kind": "viscosity"}) == {"cat_fp.kernel": 1e-4, "an_fp.kernel": 1e-4}
