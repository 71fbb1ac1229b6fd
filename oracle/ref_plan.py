"""Checker for the tile plan of the planned fused forward (csrc/fused_plan.cuh).  TEST INFRASTRUCTURE ONLY.

The plan has no counterpart in the reference: it is this library's replacement for the padding / batching step of
``build_inputs`` (train_viscosity.py:291-314; ``pad_sequences_1d`` :52-59, ``preprocess_edges_and_bonds`` :76-110).
What must hold is therefore stated against the packed CSR batch (``oracle/ref_pack.py`` is its bit-exact specification):
every molecule appears in exactly one tile, a tile's rows / entries are a faithful, tile-local copy of those molecules'
CSR rows, and rows are listed by in-degree class.  ``check_plan`` raises AssertionError otherwise and returns
statistics (tiles, fill).
"""
import numpy as np

ROWS, ECAP, MAXMOL, HEADER = 128, 336, 32, 256
TILE = np.dtype([("slot", "<u4", ROWS), ("ent", "<u4", ECAP), ("molid", "<i4", MAXMOL), ("mol_lo", "u1", MAXMOL + 4),
                 ("nm", "u1"), ("rows", "u1"), ("n_ent", "<u2"), ("pad", "u1", 24)])
assert TILE.itemsize == 2048


def parse(plan_bytes):
    """plan_bytes: uint8 numpy copy of the device buffer -> (header dict, [cation tiles, anion tiles])."""
    hdr = plan_bytes[:HEADER].view("<i4")
    n_tiles, cap, status = hdr[0:2].tolist(), hdr[2:4].tolist(), int(hdr[4])
    body = plan_bytes[HEADER:]
    tiles = []
    for t in range(2):
        off = (cap[0] if t else 0) * TILE.itemsize
        tiles.append(body[off: off + n_tiles[t] * TILE.itemsize].view(TILE))
    return {"n_tiles": n_tiles, "cap": cap, "status": status}, tiles


def _half_to_float(bits):
    return np.array(bits, dtype="<u2").view("<f2").astype(np.float64)


def check_plan(plan_bytes, host, n_pairs, atom_vocab, bond_vocab):
    """host: dict with the int32 CSR arrays mol_ptr / atom_id / row_ptr / col_src / edge_bm of the same batch."""
    hdr, tiles = parse(plan_bytes)
    assert hdr["status"] == 0, hdr
    mol_ptr, atom_id, row_ptr = host["mol_ptr"], host["atom_id"], host["row_ptr"]
    col_src, edge_bm = host["col_src"], host["edge_bm"]
    seen = np.zeros(2 * n_pairs, np.int32)
    used_rows = 0
    for tower in range(2):
        for tl in tiles[tower]:
            nm, rows = int(tl["nm"]), int(tl["rows"])
            assert 1 <= nm <= MAXMOL and rows <= ROWS
            mol_lo = tl["mol_lo"][: nm + 1].astype(int)
            assert mol_lo[0] == 0 and mol_lo[nm] == rows and (np.diff(mol_lo) >= 0).all()
            # natural rows: molecule-major copies of the CSR rows
            nat_aid, nat_deg, nat_ent = [], [], []
            for j in range(nm):
                m = int(tl["molid"][j])
                assert tower * n_pairs <= m < (tower + 1) * n_pairs
                seen[m] += 1
                a0, a1 = int(mol_ptr[m]), int(mol_ptr[m + 1])
                assert a1 - a0 == mol_lo[j + 1] - mol_lo[j]
                for at in range(a0, a1):
                    nat_aid.append(min(max(int(atom_id[at]), 0), atom_vocab - 1))
                    e0, e1 = int(row_ptr[at]), int(row_ptr[at + 1])
                    nat_deg.append(e1 - e0)
                    for e in range(e0, e1):
                        bm = int(edge_bm[e])
                        nat_ent.append((int(col_src[e]) - a0 + mol_lo[j], min(bm & 0xFFFF, bond_vocab - 1), float(bm >> 16)))
            assert int(tl["n_ent"]) == len(nat_ent) <= ECAP
            e_off = np.concatenate([[0], np.cumsum(nat_deg)]).astype(int)
            slot = tl["slot"]
            r, deg, e0, aid = slot & 127, (slot >> 7) & 31, (slot >> 12) & 1023, slot >> 22
            assert sorted(r.tolist()) == list(range(ROWS))          # a permutation of the natural rows
            key = np.minimum(deg, 7)
            assert (np.diff(key.astype(int)) >= 0).all()            # listed by in-degree class
            for pos in range(ROWS):
                rr = int(r[pos])
                if rr >= rows:
                    assert deg[pos] == 0 and aid[pos] == 0
                    continue
                assert deg[pos] == nat_deg[rr] and aid[pos] == nat_aid[rr] and e0[pos] == e_off[rr], (pos, rr)
            ent = tl["ent"][: len(nat_ent)]
            got = list(zip((ent & 0xFF).tolist(), ((ent >> 8) & 0xFF).tolist(), _half_to_float(ent >> 16).tolist()))
            assert got == nat_ent
            used_rows += rows
    assert (seen == 1).all(), "every molecule exactly once"
    n_t = sum(hdr["n_tiles"])
    return {"tiles": n_t, "fill": used_rows / max(1, n_t * ROWS)}
