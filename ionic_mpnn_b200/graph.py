"""Packed CSR graph batches -- the replacement for the reference's padded inputs.

The reference pads every ion to data-set maxima in ``build_inputs`` (train_viscosity.py:291-314:
``pad_sequences_1d`` :52-59, ``preprocess_edges_and_bonds`` :76-110) and masks padding later
(models/layers.py:74-76,114-115,163).  ``PackedGraphBatch`` holds the same information without padding;
its exact layout is specified by ``oracle/ref_pack.py`` and produced here by ``imp_pack_host`` (C++).

Three ways in, all ending in the same packer:
  * ``pack_records``  -- list of dicts in the ``src/dataset.py:15-20,51-62`` schema (ids unshifted)
  * ``pack_padded``   -- the reference's own padded input dict (drop-in for ``model.predict(x)``)
  * ``pack_flat``     -- flat ragged arrays (``FlatIons``), what ``synth_flat`` generates for benchmarks
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

I32 = np.int32


@dataclass
class FlatIons:
    """One tower as ragged int32 arrays (the flat equivalent of the record lists)."""
    atom_ptr: np.ndarray
    atom_ids: np.ndarray
    edge_ptr: np.ndarray
    edge_src: np.ndarray
    edge_dst: np.ndarray
    bond_ids: np.ndarray

    @property
    def n_ions(self):
        return len(self.atom_ptr) - 1

    def c_struct(self):
        for f in ("atom_ptr", "atom_ids", "edge_ptr", "edge_src", "edge_dst", "bond_ids"):
            a = getattr(self, f)
            assert a.dtype == I32 and a.flags.c_contiguous, f
        return _lib.Ions(self.n_ions, *(getattr(self, f).ctypes.data for f in
                                        ("atom_ptr", "atom_ids", "edge_ptr", "edge_src", "edge_dst", "bond_ids")))

    @staticmethod
    def from_ion_dicts(ions):
        atom_ptr = np.zeros(len(ions) + 1, I32)
        edge_ptr = np.zeros(len(ions) + 1, I32)
        for i, ion in enumerate(ions):
            atom_ptr[i + 1] = atom_ptr[i] + len(ion["atom_ids"])
            edge_ptr[i + 1] = edge_ptr[i] + min(len(ion["edge_indices"]), len(ion["bond_ids"]))  # zip semantics
        atom_ids = np.zeros(atom_ptr[-1], I32)
        src = np.zeros(edge_ptr[-1], I32)
        dst = np.zeros(edge_ptr[-1], I32)
        bond = np.zeros(edge_ptr[-1], I32)
        for i, ion in enumerate(ions):
            atom_ids[atom_ptr[i]:atom_ptr[i + 1]] = ion["atom_ids"]
            ne = edge_ptr[i + 1] - edge_ptr[i]
            if ne:
                e = np.asarray(ion["edge_indices"][:ne], dtype=I32).reshape(ne, 2)
                src[edge_ptr[i]:edge_ptr[i + 1]] = e[:, 0]
                dst[edge_ptr[i]:edge_ptr[i + 1]] = e[:, 1]
                bond[edge_ptr[i]:edge_ptr[i + 1]] = ion["bond_ids"][:ne]
        return FlatIons(atom_ptr, atom_ids, edge_ptr, src, dst, bond)

    def to_ion_dicts(self, lo=0, hi=None):
        """Inverse of from_ion_dicts for ions [lo, hi): the record schema of src/dataset.py:15-20."""
        hi = self.n_ions if hi is None else hi
        out = []
        for i in range(lo, hi):
            a0, a1, e0, e1 = self.atom_ptr[i], self.atom_ptr[i + 1], self.edge_ptr[i], self.edge_ptr[i + 1]
            out.append({"atom_ids": [int(v) for v in self.atom_ids[a0:a1]], "bond_ids": [int(v) for v in self.bond_ids[e0:e1]],
                        "edge_indices": [(int(a), int(b)) for a, b in zip(self.edge_src[e0:e1], self.edge_dst[e0:e1])],
                        "num_atoms": int(a1 - a0)})
        return out

    @staticmethod
    def from_padded(atom, bond, conn):
        """One tower of the reference's padded dict: atom (B,N), bond (B,E), conn (B,E,2); ids already
        shifted, edges already doubled.  Atoms kept per sample: up to the last non-zero id or the last
        atom a live edge touches (id-0 atoms inside that range stay, unpooled, as in the reference)."""
        atom = np.ascontiguousarray(atom, I32)
        bond = np.ascontiguousarray(bond, I32)
        conn = np.ascontiguousarray(conn, I32)
        B, N = atom.shape
        live = (conn[:, :, 0] > 0) & (conn[:, :, 1] > 0)
        last_id = np.where((atom > 0).any(1), N - np.argmax((atom > 0)[:, ::-1], axis=1), 0)
        last_edge = np.where(live.any(1), (conn.max(axis=2) * live).max(axis=1) + 1, 0)
        n = np.minimum(np.maximum(last_id, last_edge), N).astype(np.int64)
        atom_ptr = np.zeros(B + 1, I32)
        atom_ptr[1:] = np.cumsum(n)
        keep = np.arange(N)[None, :] < n[:, None]
        edge_ptr = np.zeros(B + 1, I32)
        edge_ptr[1:] = np.cumsum(live.sum(1))
        return FlatIons(atom_ptr, np.ascontiguousarray(atom[keep]), edge_ptr, np.ascontiguousarray(conn[:, :, 0][live]),
                        np.ascontiguousarray(conn[:, :, 1][live]), np.ascontiguousarray(bond[live]))


GRAPH_FIELDS = ("mol_ptr", "atom_id", "row_ptr", "col_src", "edge_bm", "bucket_ptr", "bucket_perm")


class PackedGraphBatch:
    """Host (numpy) arrays plus, after ``.to(device)``, their device copies (torch int32 tensors)."""

    def __init__(self, arrays, counts, temperature=None, target=None):
        self.host = arrays
        self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab = counts
        self.temperature = None if temperature is None else np.ascontiguousarray(temperature, np.float32).reshape(-1)
        self.target = None if target is None else np.ascontiguousarray(target, np.float32).reshape(-1)
        mp = self.host.get("mol_ptr")
        self.max_mol_atoms = int(np.diff(mp).max()) if mp is not None and len(mp) > 1 else 0
        # envelope of the tile plan (imp_fused_plan): in-degree <= 31, <= 336 unique entries per molecule
        rp = self.host.get("row_ptr")
        self.max_in_degree = int(np.diff(rp).max()) if rp is not None and len(rp) > 1 else 0
        self.max_mol_entries = int(np.diff(rp[mp]).max()) if rp is not None and mp is not None and len(mp) > 1 else 0
        self.dev = None
        self.dev_T = None
        self.dev_y = None
        self.device = None
        self.symmetric = None  # True / False / None (unknown: checked before the first training use)

    # -- device side ------------------------------------------------------------------------
    def to(self, device, non_blocking=False, pinned=False):
        import torch

        self.device = torch.device(device)
        self.dev = {}
        for k in GRAPH_FIELDS:
            t = torch.from_numpy(self.host[k])
            if pinned:
                t = t.pin_memory()
            self.dev[k] = t.to(self.device, non_blocking=non_blocking)
        if self.temperature is not None:
            self.dev_T = torch.from_numpy(self.temperature).to(self.device, non_blocking=non_blocking)
        if self.target is not None:
            self.dev_y = torch.from_numpy(self.target).to(self.device, non_blocking=non_blocking)
        return self

    def pin(self):
        """Page-locks the host arrays (once) so that H2D copies can run asynchronously on a copy stream."""
        import torch

        if getattr(self, "pinned", None) is None:
            self.pinned = {k: torch.from_numpy(self.host[k]).pin_memory() for k in GRAPH_FIELDS}
            self.pinned_T = None if self.temperature is None else torch.from_numpy(self.temperature).pin_memory()
        return self

    def compact(self):
        """Builds (once) the compact input feed of the fused forward (include/imp_b200.h, imp_compact_graph_t): 16-bit atom
        words, 32-bit entry words with molecule-local sources, per-molecule entry offsets.  Raises if the batch does
        not fit the format (ids >= 256, in-degree > 255, molecule > 256 atoms, multiplicity > 255)."""
        if getattr(self, "chost", None) is not None:
            return self
        h = self.host
        mol_ptr, row_ptr = h["mol_ptr"].astype(np.int64), h["row_ptr"].astype(np.int64)
        deg = np.diff(row_ptr)
        n_mols = len(mol_ptr) - 1
        atom_mol_base = np.repeat(mol_ptr[:-1], np.diff(mol_ptr))
        entry_dst = np.repeat(np.arange(self.n_atoms, dtype=np.int64), deg)
        src_local = h["col_src"].astype(np.int64) - atom_mol_base[entry_dst] if self.n_unique else np.zeros(0, np.int64)
        bond = h["edge_bm"].astype(np.int64) & 0xFFFF
        mult = h["edge_bm"].astype(np.int64) >> 16
        ok = (self.n_atoms == 0 or (h["atom_id"].min() >= 0 and h["atom_id"].max() < 256 and deg.max() <= 255)) and \
             (self.n_unique == 0 or (src_local.min() >= 0 and src_local.max() < 256 and bond.max() < 256 and mult.max() < 256))
        if not ok:
            raise _lib.ImpError("batch does not fit the compact feed (8-bit ids / degrees / local sources)")
        self.chost = {
            "mol_ptr": h["mol_ptr"],
            "mol_eptr": np.ascontiguousarray(row_ptr[mol_ptr].astype(I32)) if n_mols >= 0 else np.zeros(1, I32),
            "atom_w": np.ascontiguousarray((h["atom_id"].astype(np.int64) | (deg << 8)).astype(np.uint16)),
            "edge_w": np.ascontiguousarray((src_local | (bond << 8) | (mult << 16)).astype(np.uint32)),
        }
        # narrow form (imp_fused_plan_compact16): 16-bit entry words when sources fit 7 bits and multiplicities are 1 or 2
        self.narrow_ok = bool(self.n_unique == 0 or (src_local.max() < 128 and mult.min() >= 1 and mult.max() <= 2))
        if self.narrow_ok:
            self.chost["edge_h"] = np.ascontiguousarray((src_local | (bond << 7) | ((mult - 1) << 15)).astype(np.uint16))
        return self

    def nbytes_compact(self):
        self.compact()
        return sum(v.nbytes for v in self.chost.values()) + (0 if self.temperature is None else self.temperature.nbytes)

    def pin_compact(self):
        import torch

        self.compact()
        if getattr(self, "cpinned", None) is None:
            views = {"mol_ptr": self.chost["mol_ptr"], "mol_eptr": self.chost["mol_eptr"],
                     "atom_w": self.chost["atom_w"].view(np.int16), "edge_w": self.chost["edge_w"].view(np.int32)}
            if self.narrow_ok:
                views["edge_h"] = self.chost["edge_h"].view(np.int16)
            self.cpinned = {k: torch.from_numpy(v).pin_memory() for k, v in views.items()}
            self.pinned_T = None if self.temperature is None else torch.from_numpy(self.temperature).pin_memory()
        return self

    def to_compact(self, device, narrow=False):
        """Device copy of the compact feed only (what imp_mpnn_forward_fused_compact / imp_fused_plan read); ``narrow``: the
        16-bit entry words of imp_fused_plan_compact16 (the planned forward only)."""
        import torch

        self.pin_compact()
        if narrow and not self.narrow_ok:
            raise _lib.ImpError("batch does not fit the narrow compact feed (molecule-local sources < 128, multiplicity 1 or 2)")
        slot = DeviceSlot(torch.device(device), COMPACT16_FIELDS if narrow else COMPACT_FIELDS)
        slot.load(self, torch.cuda.current_stream())
        return slot

    def c_struct(self):
        if self.dev is None:
            raise _lib.ImpError("PackedGraphBatch is not on a device: call .to('cuda') first")
        return _lib.Graph(self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab,
                          *(self.dev[k].data_ptr() for k in GRAPH_FIELDS))

    def nbytes(self):
        n = sum(self.host[k].nbytes for k in GRAPH_FIELDS)
        return n + (0 if self.temperature is None else self.temperature.nbytes)


FUSED_FIELDS = ("mol_ptr", "atom_id", "row_ptr", "col_src", "edge_bm")  # what imp_mpnn_forward_fused reads
COMPACT_FIELDS = ("mol_ptr", "mol_eptr", "atom_w", "edge_w")              # what imp_mpnn_forward_fused_compact reads
COMPACT16_FIELDS = ("mol_ptr", "mol_eptr", "atom_w", "edge_h")            # what imp_fused_plan_compact16 reads


class DeviceSlot:
    """Device staging buffers for one in-flight chunk of a streamed prediction (see MPNNModel.predict_stream).
    Quacks like a PackedGraphBatch on the device side."""

    def __init__(self, device, fields):
        self.device, self.fields = device, fields
        self.dev = {}
        self.dev_T = None
        self.dev_y = None
        self.cap = {}

    @property
    def is_compact(self):
        return self.fields in (COMPACT_FIELDS, COMPACT16_FIELDS)

    @property
    def is_narrow(self):
        return self.fields == COMPACT16_FIELDS

    def load(self, chunk, stream):
        """Enqueues the H2D copies of ``chunk`` (pinned) on ``stream``; returns the bytes copied."""
        import torch

        n = 0
        pinned = chunk.cpinned if self.is_compact else chunk.pinned
        with torch.cuda.stream(stream):
            for k in self.fields:
                src = pinned[k]
                if self.cap.get(k, 0) < src.numel():
                    self.cap[k] = int(src.numel() * 1.05) + 16
                    self.dev[k + "_buf"] = torch.empty(self.cap[k], dtype=src.dtype, device=self.device)
                self.dev[k] = self.dev[k + "_buf"][: src.numel()]
                self.dev[k].copy_(src, non_blocking=True)
                n += src.numel() * src.element_size()
            if not self.is_compact:
                for k in GRAPH_FIELDS:  # fields the fused path never reads
                    if k not in self.fields:
                        self.dev[k] = self.dev[self.fields[0]]
            if chunk.pinned_T is not None:
                if self.cap.get("T", 0) < chunk.pinned_T.numel():
                    self.cap["T"] = int(chunk.pinned_T.numel() * 1.05) + 16
                    self._T_buf = torch.empty(self.cap["T"], dtype=torch.float32, device=self.device)
                self.dev_T = self._T_buf[: chunk.pinned_T.numel()]
                self.dev_T.copy_(chunk.pinned_T, non_blocking=True)
                n += chunk.pinned_T.numel() * 4
        for a in ("n_pairs", "n_atoms", "n_cat_atoms", "n_unique", "n_edges", "bond_vocab", "max_mol_atoms"):
            setattr(self, a, getattr(chunk, a))
        for a in ("max_in_degree", "max_mol_entries"):
            setattr(self, a, getattr(chunk, a, None))
        return n

    def compact_struct(self):
        return _lib.CompactGraph(self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab,
                                 *(self.dev[k].data_ptr() for k in self.fields))

    def c_struct(self):
        return _lib.Graph(self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab,
                          *(self.dev[k].data_ptr() for k in GRAPH_FIELDS))

    def to(self, device):
        return self


def pack_flat(cation: FlatIons, anion: FlatIons, bond_vocab_size, max_edges=None, double_edges=True, shift_ids=True,
              temperature=None, target=None, n_threads=0):
    assert cation.n_ions == anion.n_ions
    P = cation.n_ions
    N = int(cation.atom_ptr[-1]) + int(anion.atom_ptr[-1])
    cap = (int(cation.edge_ptr[-1]) + int(anion.edge_ptr[-1])) * (2 if double_edges else 1)
    arrays = {
        "mol_ptr": np.zeros(2 * P + 1, I32), "atom_id": np.zeros(N, I32), "row_ptr": np.zeros(N + 1, I32),
        "col_src": np.zeros(cap, I32), "edge_bm": np.zeros(cap, I32),
        "bucket_ptr": np.zeros(2 * bond_vocab_size + 1, I32), "bucket_perm": np.zeros(cap, I32),
    }
    g = _lib.Graph(0, 0, 0, 0, 0, 0, *(arrays[k].ctypes.data for k in GRAPH_FIELDS))
    flags = (_lib.PACK_DOUBLE_EDGES if double_edges else 0) | (_lib.PACK_SHIFT_IDS if shift_ids else 0)
    cs, an = cation.c_struct(), anion.c_struct()
    _lib.call("imp_pack_host", C.byref(cs), C.byref(an), bond_vocab_size, -1 if max_edges is None else int(max_edges),
              flags, cap, C.byref(g), n_threads)
    for k in ("col_src", "edge_bm", "bucket_perm"):
        arrays[k] = arrays[k][: g.n_unique]
    counts = (g.n_pairs, g.n_atoms, g.n_cat_atoms, g.n_unique, g.n_edges, g.bond_vocab)
    b = PackedGraphBatch(arrays, counts, temperature, target)
    # reverse-doubled and untruncated: every live entry has its mirror (train_viscosity.py:87-91).  Anything else is checked
    # on the host when the batch is first used for training (train.entries_are_symmetric).
    b.symmetric = True if (double_edges and max_edges is None) else None
    return b


def pack_records(records, bond_vocab_size, max_edges=None, label=None):
    cat = FlatIons.from_ion_dicts([r["cation"] for r in records])
    an = FlatIons.from_ion_dicts([r["anion"] for r in records])
    T = np.array([r["T"] for r in records], np.float32) if records and "T" in records[0] else None
    y = np.array([r[label] for r in records], np.float32) if label else None
    return pack_flat(cat, an, bond_vocab_size, max_edges=max_edges, temperature=T, target=y)


def pack_padded(x, bond_vocab_size):
    """``x``: the reference's input dict (train_viscosity.py:306-314; ``temperature`` optional)."""
    cat = FlatIons.from_padded(x["cat_atom"], x["cat_bond"], x["cat_connectivity"])
    an = FlatIons.from_padded(x["an_atom"], x["an_bond"], x["an_connectivity"])
    T = np.asarray(x["temperature"], np.float32).reshape(-1) if "temperature" in x else None
    return pack_flat(cat, an, bond_vocab_size, double_edges=False, shift_ids=False, temperature=T)


def synth_flat(n_ions, seed, n_min=10, n_max=40, atom_types=123, bond_types=71, skewed=False):
    """Benchmark-sized synthetic tower via ``imp_synth_ions`` (same recipe as synth.make_ion)."""
    na, ne = C.c_int64(0), C.c_int64(0)
    _lib.call("imp_synth_ions", seed, n_ions, n_min, n_max, atom_types, bond_types, int(skewed), None, None, None, None,
              None, None, C.byref(na), C.byref(ne))
    ions = FlatIons(np.zeros(n_ions + 1, I32), np.zeros(na.value, I32), np.zeros(n_ions + 1, I32),
                    np.zeros(ne.value, I32), np.zeros(ne.value, I32), np.zeros(ne.value, I32))
    _lib.call("imp_synth_ions", seed, n_ions, n_min, n_max, atom_types, bond_types, int(skewed), ions.atom_ptr.ctypes.data,
              ions.atom_ids.ctypes.data, ions.edge_ptr.ctypes.data, ions.edge_src.ctypes.data, ions.edge_dst.ctypes.data,
              ions.bond_ids.ctypes.data, C.byref(na), C.byref(ne))
    return ions


def synth_batch(n_pairs, seed, n_min=10, n_max=40, skewed=False, bond_vocab_size=72, with_temperature=True):
    cat = synth_flat(n_pairs, 2 * seed + 1, n_min, n_max, skewed=skewed)
    an = synth_flat(n_pairs, 2 * seed + 2, n_min, n_max, skewed=skewed)
    T = None
    if with_temperature:
        T = np.random.default_rng(seed).uniform(273.15, 373.15, size=n_pairs).astype(np.float32)
    return pack_flat(cat, an, bond_vocab_size, temperature=T), cat, an


ION_FIELDS = ("atom_ptr", "atom_ids", "edge_ptr", "edge_src", "edge_dst", "bond_ids")


class DevicePackedBatch:
    """A batch packed ON the device by imp_pack_device: int32 CSR arrays + the compact feed, no host copy and no
    bond buckets (forward only: fused kernel, fp32 staged kernels with fused messages).  Quacks like a PackedGraphBatch
    on the device side; ``as_compact()`` returns the view that routes through imp_mpnn_forward_fused_compact."""

    is_compact = False

    def __init__(self, dev, counts, max_mol_atoms, dev_T=None):
        self.dev, self.dev_T, self.dev_y = dev, dev_T, None
        self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab = counts
        self.max_mol_atoms = max_mol_atoms
        self.max_in_degree = self.max_mol_entries = None  # not known on the host: the tile plan's status word reports them

    def c_struct(self):
        ptr = lambda k: self.dev[k].data_ptr() if k in self.dev else 0  # noqa: E731
        return _lib.Graph(self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab,
                          *(ptr(k) for k in GRAPH_FIELDS))

    def compact_struct(self):
        return _lib.CompactGraph(self.n_pairs, self.n_atoms, self.n_cat_atoms, self.n_unique, self.n_edges, self.bond_vocab,
                                 *(self.dev[k].data_ptr() for k in COMPACT_FIELDS))

    def as_compact(self):
        import copy

        c = copy.copy(self)
        c.is_compact = True
        return c

    def to(self, device):
        return self


def upload_ions(cation: FlatIons, anion: FlatIons, device="cuda", pin=False):
    """H2D copy of the flat ragged ion arrays of both towers (optionally through pinned staging buffers)."""
    import torch

    dev = torch.device(device)
    out = {}
    for t, ions in (("cat", cation), ("an", anion)):
        out[t] = {}
        for f in ION_FIELDS:
            src = torch.from_numpy(getattr(ions, f))
            if pin:
                src = src.pin_memory()
            out[t][f] = src.to(dev, non_blocking=True)
    return out


def pack_flat_device(cation: FlatIons, anion: FlatIons, bond_vocab_size, device="cuda", max_edges=None, double_edges=True,
                     shift_ids=True, temperature=None, compact=True, stream=None, uploaded=None):
    """Uploads the flat ragged ion arrays (unless ``uploaded`` = upload_ions(...) is given) and builds the packed batch
    on the GPU (imp_pack_device).  One host synchronisation at the end to read the counts / status back."""
    import torch

    assert cation.n_ions == anion.n_ions
    dev = torch.device(device)
    P = cation.n_ions
    n_cat, n_an = int(cation.atom_ptr[-1]), int(anion.atom_ptr[-1])
    N = n_cat + n_an
    cap = (int(cation.edge_ptr[-1]) + int(anion.edge_ptr[-1])) * (2 if double_edges else 1)
    up = uploaded if uploaded is not None else upload_ions(cation, anion, dev)
    cs = _lib.Ions(P, *(up["cat"][f].data_ptr() for f in ION_FIELDS))
    an = _lib.Ions(P, *(up["an"][f].data_ptr() for f in ION_FIELDS))
    i32 = dict(dtype=torch.int32, device=dev)
    out = {"mol_ptr": torch.empty(2 * P + 1, **i32), "atom_id": torch.empty(max(N, 1), **i32),
           "row_ptr": torch.empty(N + 1, **i32), "col_src": torch.empty(max(cap, 1), **i32),
           "edge_bm": torch.empty(max(cap, 1), **i32)}
    if compact:
        out["mol_eptr"] = torch.empty(2 * P + 1, **i32)
        out["atom_w"] = torch.empty(max(N, 1), dtype=torch.int16, device=dev)
        out["edge_w"] = torch.empty(max(cap, 1), **i32)
    counts = torch.zeros(4, **i32)
    ws = torch.empty(int(_lib.load().imp_pack_device_workspace_bytes(P)), dtype=torch.uint8, device=dev)
    g = _lib.Graph(0, 0, 0, 0, 0, 0, out["mol_ptr"].data_ptr(), out["atom_id"].data_ptr(), out["row_ptr"].data_ptr(),
                   out["col_src"].data_ptr(), out["edge_bm"].data_ptr(), 0, 0)
    flags = (_lib.PACK_DOUBLE_EDGES if double_edges else 0) | (_lib.PACK_SHIFT_IDS if shift_ids else 0)
    st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
    opt = lambda k: out[k].data_ptr() if compact else None  # noqa: E731
    _lib.call("imp_pack_device", C.byref(cs), C.byref(an), n_cat, N, bond_vocab_size, -1 if max_edges is None else int(max_edges),
              flags, cap, C.byref(g), opt("mol_eptr"), opt("atom_w"), opt("edge_w"), counts.data_ptr(), ws.data_ptr(), st)
    c = counts.cpu().numpy()  # synchronises
    if c[1] != 0:
        raise _lib.ImpError(f"imp_pack_device: device packer reported error {int(c[1])} "
                            "(-3 index out of range, -4 molecule too large for the device packer / capacity, -2 compact range)")
    n_unique = int(c[0])
    n_edges = int(np.uint32(c[2])) | (int(np.uint32(c[3])) << 32)
    for k in ("col_src", "edge_bm", "edge_w"):
        if k in out:
            out[k] = out[k][: max(n_unique, 1)]
    sizes = np.concatenate([np.diff(cation.atom_ptr), np.diff(anion.atom_ptr)]) if P else np.zeros(0, I32)
    dev_T = None if temperature is None else torch.from_numpy(np.ascontiguousarray(temperature, np.float32)).to(dev)
    return DevicePackedBatch(out, (P, N, n_cat, n_unique, n_edges, bond_vocab_size), int(sizes.max()) if P else 0, dev_T)
