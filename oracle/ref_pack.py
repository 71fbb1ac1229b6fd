"""TEST INFRASTRUCTURE ONLY -- plain-Python definition of the packed CSR graph batch.

The reference pads every ion to the data-set maxima (train_viscosity.py:288-314); the B200
path replaces that by one packed batch.  This file is the bit-exact specification the C++ host
packer (``imp_pack_host``) and the CUDA packer (``imp_pack_device``) are tested against; it is
derived from the reference semantics as follows (paths relative to /root/reference):

* ids are shifted ``+1``, edge indices are not            train_viscosity.py:255-262
* each featurize entry ``(s, t, b)`` yields ``(s, t, b)`` and ``(t, s, b)``, list truncated to
  ``2 * max_edges`` entries when ``max_edges`` is given     train_viscosity.py:87-105
* an entry is live iff ``s > 0 and t > 0``                  models/layers.py:114-115 (message mask),
                                                            models/layers.py:74-76 (Reduce drops tgt==0)
* all atoms of an ion (id > 0 after the shift) are pooled   models/layers.py:161-164

Layout (all int32):
  mol_ptr[2P+1]     atom offsets, molecule m < P is the cation of pair m, m >= P the anion of pair m-P
  atom_id[N]        shifted vocabulary ids (1..V_a-1)
  row_ptr[N+1]      CSR over DESTINATION atoms
  col_src[Eu]       global source atom of each unique live entry, rows sorted by (bond, src)
  edge_bm[Eu]       bond id (shifted, low 16 bits) | multiplicity << 16
  bucket_ptr[2*Vb+1], bucket_perm[Eu]   unique entries grouped by (tower, bond id), stable in CSR order
  n_edges           sum of multiplicities == number of live padded-array entries (the "E" of SURVEY 8d)
"""
from __future__ import annotations

import numpy as np


def live_entries(ion, max_edges=None):
    """Doubled, truncated, masked entry list of one ion: [(src, tgt, shifted_bond), ...] in the
    reference's order."""
    out = []
    for (s, t), b in zip(ion["edge_indices"], ion["bond_ids"]):
        out.append((int(s), int(t), int(b) + 1))
        out.append((int(t), int(s), int(b) + 1))
    if max_edges is not None:
        out = out[: 2 * max_edges]
    return [(s, t, b) for (s, t, b) in out if s > 0 and t > 0]


def pack_records(records, bond_vocab_size, max_edges=None):
    P = len(records)
    mols = [r["cation"] for r in records] + [r["anion"] for r in records]
    mol_ptr = np.zeros(2 * P + 1, dtype=np.int64)
    for m, ion in enumerate(mols):
        mol_ptr[m + 1] = mol_ptr[m] + len(ion["atom_ids"])
    N = int(mol_ptr[-1])
    atom_id = np.zeros(N, dtype=np.int32)
    triples = []  # (dst, bond, src) global
    for m, ion in enumerate(mols):
        base, n = int(mol_ptr[m]), len(ion["atom_ids"])
        atom_id[base : base + n] = np.asarray(ion["atom_ids"], dtype=np.int32) + 1
        for s, t, b in live_entries(ion, max_edges):
            if not (s < n and t < n):
                raise ValueError(f"edge ({s},{t}) out of range for ion with {n} atoms")
            if not (0 < b < bond_vocab_size):
                raise ValueError(f"bond id {b} outside vocabulary of {bond_vocab_size}")
            triples.append((base + t, b, base + s))
    triples.sort()
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    col_src, edge_bm = [], []
    i = 0
    while i < len(triples):
        j = i
        while j < len(triples) and triples[j] == triples[i]:
            j += 1
        dst, b, src = triples[i]
        mult = j - i
        if mult >= 1 << 15:
            raise ValueError("multiplicity overflow")
        col_src.append(src)
        edge_bm.append(b | (mult << 16))
        row_ptr[dst + 1] += 1
        i = j
    row_ptr = np.cumsum(row_ptr)
    col_src = np.asarray(col_src, dtype=np.int32)
    edge_bm = np.asarray(edge_bm, dtype=np.int32)
    Eu = len(col_src)
    n_cat = int(mol_ptr[P])
    # destination of each unique entry (for the tower of the bucket key)
    dst_of = np.repeat(np.arange(N, dtype=np.int64), np.diff(row_ptr)) if N else np.zeros(0, np.int64)
    key = (dst_of >= n_cat).astype(np.int64) * bond_vocab_size + (edge_bm & 0xFFFF)
    bucket_perm = np.argsort(key, kind="stable").astype(np.int32)
    bucket_ptr = np.zeros(2 * bond_vocab_size + 1, dtype=np.int64)
    np.add.at(bucket_ptr, key + 1, 1)
    bucket_ptr = np.cumsum(bucket_ptr)
    return {
        "n_pairs": P,
        "n_atoms": N,
        "n_cat_atoms": n_cat,
        "n_unique": Eu,
        "n_edges": int(sum(e >> 16 for e in edge_bm.tolist())),
        "mol_ptr": mol_ptr.astype(np.int32),
        "atom_id": atom_id,
        "row_ptr": row_ptr.astype(np.int32),
        "col_src": col_src,
        "edge_bm": edge_bm,
        "bucket_ptr": bucket_ptr.astype(np.int32),
        "bucket_perm": bucket_perm,
    }


def padded_live_multiset(x, prefix):
    """Multiset {(sample, tgt, bond, src): count} of live entries of one tower of the reference's
    padded input dict -- what BondMatrixMessage+Reduce actually sum (models/layers.py:74-76,114-115)."""
    conn, bond = x[f"{prefix}_connectivity"], x[f"{prefix}_bond"]
    out = {}
    for i in range(conn.shape[0]):
        for e in range(conn.shape[1]):
            s, t = int(conn[i, e, 0]), int(conn[i, e, 1])
            if s > 0 and t > 0:
                k = (i, t, int(bond[i, e]), s)
                out[k] = out.get(k, 0) + 1
    return out


def packed_live_multiset(pk, tower):
    """Same multiset, reconstructed from the packed batch."""
    P = pk["n_pairs"]
    out = {}
    for m in range(tower * P, (tower + 1) * P):
        a0, a1 = int(pk["mol_ptr"][m]), int(pk["mol_ptr"][m + 1])
        for v in range(a0, a1):
            for e in range(int(pk["row_ptr"][v]), int(pk["row_ptr"][v + 1])):
                bm = int(pk["edge_bm"][e])
                k = (m - tower * P, v - a0, bm & 0xFFFF, int(pk["col_src"][e]) - a0)
                out[k] = out.get(k, 0) + (bm >> 16)
    return out
