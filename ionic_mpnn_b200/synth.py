"""Deterministic synthetic ion-pair records in the reference's record schema.

The reference ships no data (its ``data/`` is git-ignored), so parity and benchmark inputs are
synthetic.  Records have exactly the layout ``src/dataset.py:15-20,51-62`` writes and follow the
``src/featurize.py:60-63`` convention: every chemical bond contributes the two consecutive
entries ``(a, b)``, ``(b, a)`` that share one bond id, ids are 0-based vocabulary indices (the
``+1`` shift happens later, ``train_viscosity.py:255-262``).

Generator: per ion ``n ~ U{n_min..n_max}``, a random tree with
``parent(i) ~ U{max(0,i-3)..i-1}`` under a degree cap of 4, plus ``r ~ U{0..2}`` ring-closing
bonds between distinct non-adjacent atoms; atom type ``~ U{0..122}``; bond type uniform over
``{0..70}`` or Zipf(1.2)-skewed; ``T ~ U[273.15, 373.15]``; labels ``log_eta ~ N(2,1)``,
``mp ~ N(330,60)``.

This pure-Python generator is for parity-sized sets; ``imp_synth_pairs`` in the C-ABI library is
the same recipe with a splitmix64 stream for benchmark-sized sets (millions of pairs).
"""
from __future__ import annotations

import numpy as np

ATOM_TYPES = 123  # README.md:175 (vocab.pkl atom_vocab_size)
BOND_TYPES = 71   # README.md:176


def _zipf_table(n, a=1.2):
    w = 1.0 / np.arange(1, n + 1) ** a
    return np.cumsum(w / w.sum())


def make_ion(rng, n_min=10, n_max=40, skewed=False, atom_types=ATOM_TYPES, bond_types=BOND_TYPES):
    n = int(rng.integers(n_min, n_max + 1))
    deg = [0] * n
    adj = set()
    bonds = []
    for i in range(1, n):
        cand = [j for j in range(max(0, i - 3), i) if deg[j] < 4]
        if not cand:
            cand = [j for j in range(i) if deg[j] < 4]
        j = cand[int(rng.integers(len(cand)))]
        bonds.append((j, i))
        adj.add((j, i))
        deg[i] += 1
        deg[j] += 1
    for _ in range(int(rng.integers(0, 3))):
        a, b = int(rng.integers(n)), int(rng.integers(n))
        lo, hi = min(a, b), max(a, b)
        if lo != hi and (lo, hi) not in adj and deg[lo] < 4 and deg[hi] < 4:
            bonds.append((lo, hi))
            adj.add((lo, hi))
            deg[lo] += 1
            deg[hi] += 1
    if skewed:
        cdf = _zipf_table(bond_types)
        btypes = np.searchsorted(cdf, rng.random(len(bonds))).clip(0, bond_types - 1)
    else:
        btypes = rng.integers(0, bond_types, size=len(bonds))
    edge_indices, bond_ids = [], []
    for (a, b), t in zip(bonds, btypes):
        edge_indices += [(a, b), (b, a)]
        bond_ids += [int(t), int(t)]
    return {
        "atom_ids": [int(v) for v in rng.integers(0, atom_types, size=n)],
        "bond_ids": bond_ids,
        "edge_indices": edge_indices,
        "num_atoms": n,
    }


def make_records(n_pairs, seed=0, n_min=10, n_max=40, skewed=False, label="log_eta",
                 atom_types=ATOM_TYPES, bond_types=BOND_TYPES):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n_pairs):
        r = {
            "pair_id": i,
            "cation": make_ion(rng, n_min, n_max, skewed, atom_types, bond_types),
            "anion": make_ion(rng, n_min, n_max, skewed, atom_types, bond_types),
        }
        if label == "log_eta":
            r["T"] = float(np.float32(rng.uniform(273.15, 373.15)))
            r["log_eta"] = float(rng.normal(2.0, 1.0))
        else:
            r["mp"] = float(rng.normal(330.0, 60.0))
        recs.append(r)
    return recs
