"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's input construction.

Follows (reference paths relative to /root/reference):
  * ``pad_sequences_1d``            train_viscosity.py:52-59
  * ``preprocess_edges_and_bonds``  train_viscosity.py:76-110 (copies: train_melting_point.py:64-97,
                                    utils/mp_utils.py:17-44)
  * the ``+1`` id shifts            train_viscosity.py:255-262
  * ``build_inputs``                train_viscosity.py:291-314 / train_melting_point.py:253-273
  * record schema                   src/dataset.py:15-20,51-62

Quirks reproduced on purpose: every featurize edge entry is emitted
again together with its reverse (multiplicity 2 per directed edge), ``edge_indices`` are NOT
shifted, and padding edges are ``[0, 0]`` with bond id 0.
"""
from __future__ import annotations

import numpy as np


def pad_sequences_1d(seq_list, max_len, pad_val=0):
    out = np.full((len(seq_list), max_len), pad_val, dtype=np.int32)
    for i, s in enumerate(seq_list):
        if len(s) > max_len:  # the reference would build a ragged array and fail
            raise ValueError("sequence longer than max_len")
        out[i, : len(s)] = s
    return out


def preprocess_edges_and_bonds(edge_list, bond_list, max_edges):
    """Returns ``conn (B, 2*max_edges, 2) int32`` and ``bond (B, 2*max_edges) int32``."""
    max_len = 2 * max_edges
    conn = np.zeros((len(edge_list), max_len, 2), dtype=np.int32)
    bond = np.zeros((len(edge_list), max_len), dtype=np.int32)
    for i, (edges, bonds) in enumerate(zip(edge_list, bond_list)):
        k = 0
        for (src, tgt), b in zip(edges, bonds):  # zip: stops at the shorter list, like the reference
            for s, t in ((src, tgt), (tgt, src)):
                if k < max_len:  # "e[:max_len]" truncation
                    conn[i, k, 0], conn[i, k, 1], bond[i, k] = s, t, b
                k += 1
    return conn, bond


def build_inputs(records, max_atoms=None, max_edges=None, with_temperature=True):
    """records: list of dicts in the src/dataset.py schema.  Returns the dict of padded numpy
    arrays the reference feeds to ``model.predict`` (keys as train_viscosity.py:306-314)."""
    cat_atoms = [[a + 1 for a in r["cation"]["atom_ids"]] for r in records]
    cat_bonds = [[b + 1 for b in r["cation"]["bond_ids"]] for r in records]
    cat_edges = [r["cation"]["edge_indices"] for r in records]
    an_atoms = [[a + 1 for a in r["anion"]["atom_ids"]] for r in records]
    an_bonds = [[b + 1 for b in r["anion"]["bond_ids"]] for r in records]
    an_edges = [r["anion"]["edge_indices"] for r in records]
    if max_atoms is None:
        max_atoms = max(max(map(len, cat_atoms)), max(map(len, an_atoms)))
    if max_edges is None:
        max_edges = max(max(map(len, cat_edges)), max(map(len, an_edges)))
    ce, cb = preprocess_edges_and_bonds(cat_edges, cat_bonds, max_edges)
    ae, ab = preprocess_edges_and_bonds(an_edges, an_bonds, max_edges)
    x = {
        "cat_atom": pad_sequences_1d(cat_atoms, max_atoms),
        "cat_bond": cb,
        "cat_connectivity": ce,
        "an_atom": pad_sequences_1d(an_atoms, max_atoms),
        "an_bond": ab,
        "an_connectivity": ae,
    }
    if with_temperature:
        x["temperature"] = np.array([r["T"] for r in records], np.float32)[:, None]
    return x
