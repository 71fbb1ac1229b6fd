"""Feature tuples -> vocabulary -> integer-id records: the host stage between the reference's RDKit featuriser and the packed
graph batch (the ids it writes are what ``graph.pack_records`` / ``FlatIons.from_ion_dicts`` read).

Mirrors, on plain Python containers,

* ``src/featurize.py:8-29,32-76``   ``smiles_to_graph``: RDKit perception (symbol, formal charge, total H count after AddHs,
  aromaticity, hybridisation; bond type, conjugation, ring membership).  RDKit is not part of this image: the function
  delegates to it when it can be imported and raises ImportError otherwise -- there is no re-implementation of RDKit's
  chemistry here, and no parity claim for it;
* ``src/build_vocab.py:18-66``       ``build_vocab``: sorted unique feature tuples -> 0-based ids (the ``+1`` shift that reserves
  id 0 for padding happens later, ``train_viscosity.py:255-262`` = ``graph.pack_records``);
* ``src/dataset.py:5-20,23-79``      ``convert_graph_to_ids`` / ``process_records``: tuples -> ids, records with a feature outside
  the vocabulary are skipped and reported, labels (T, log_eta / mp) carried over.
"""
from __future__ import annotations


ATOM_FIELDS = ("GetSymbol", "GetFormalCharge", "GetTotalNumHs", "GetIsAromatic", "GetHybridization")  # src/featurize.py:13-17
BOND_FIELDS = ("GetBondType", "GetIsConjugated", "IsInRing")                                            # src/featurize.py:25-27


def _atom_tuple(atom):
    sym, charge, n_h, arom, hyb = (getattr(atom, f)() for f in ATOM_FIELDS)
    return (sym, charge, n_h, int(arom), str(hyb))


def _bond_tuple(bond):
    kind, conj, ring = (getattr(bond, f)() for f in BOND_FIELDS)
    return (str(kind), conj, ring)


def smiles_to_graph(smiles):
    """The graph dict of src/featurize.py:32-76 for one SMILES string (needs RDKit: explicit hydrogens added, one feature
    tuple per atom; every bond contributes the two consecutive entries (a, b), (b, a) that share one feature tuple -- the
    convention ``ionic_mpnn_b200.synth`` and the packers rely on)."""
    try:
        from rdkit import Chem
    except ImportError as e:  # pragma: no cover - RDKit is absent from this image
        raise ImportError("smiles_to_graph needs RDKit (src/featurize.py delegates the chemistry to it); feed feature-tuple graphs "
                          "or integer-id records instead") from e
    mol = Chem.MolFromSmiles(smiles)
    if mol is None:
        raise ValueError(f"invalid SMILES string: {smiles}")
    mol = Chem.AddHs(mol)
    ends = [(b.GetBeginAtomIdx(), b.GetEndAtomIdx(), _bond_tuple(b)) for b in mol.GetBonds()]
    atoms = [_atom_tuple(a) for a in mol.GetAtoms()]
    return {"smiles": smiles, "atom_features": atoms,
            "bond_features": [f for _, _, f in ends for _ in (0, 1)],
            "edge_indices": [e for u, v, _ in ends for e in ((u, v), (v, u))],
            "num_atoms": len(atoms)}


def build_vocab(*datasets):
    """src/build_vocab.py:18-66 over any number of graph-record lists (the reference reads the viscosity and the melting-point
    file): ids follow the SORTED order of the unique feature tuples."""
    atom_set, bond_set = set(), set()
    for data in datasets:
        for rec in data:
            for ion in ("cation_graph", "anion_graph"):
                atom_set.update(rec[ion]["atom_features"])
                bond_set.update(rec[ion]["bond_features"])
    atom_vocab = {feat: idx for idx, feat in enumerate(sorted(atom_set))}
    bond_vocab = {feat: idx for idx, feat in enumerate(sorted(bond_set))}
    return {"atom_vocab": atom_vocab, "bond_vocab": bond_vocab, "atom_vocab_size": len(atom_vocab), "bond_vocab_size": len(bond_vocab)}


def convert_graph_to_ids(graph, atom_vocab, bond_vocab):
    """src/dataset.py:5-20.  KeyError for a feature outside the vocabulary."""
    atom_ids = [atom_vocab[f] for f in graph["atom_features"]]
    bond_ids = [bond_vocab[f] for f in graph["bond_features"]]
    return {"atom_ids": atom_ids, "bond_ids": bond_ids, "edge_indices": graph["edge_indices"], "num_atoms": len(atom_ids)}


def process_records(data, vocab):
    """src/dataset.py:23-79 without the files: returns (id records, skipped) where skipped lists the pair ids whose graphs hold
    a feature that is not in the vocabulary."""
    atom_vocab, bond_vocab = vocab["atom_vocab"], vocab["bond_vocab"]
    out, skipped = [], []
    for rec in data:
        try:
            new = {"pair_id": rec["pair_id"], "cation": convert_graph_to_ids(rec["cation_graph"], atom_vocab, bond_vocab),
                   "anion": convert_graph_to_ids(rec["anion_graph"], atom_vocab, bond_vocab)}
        except KeyError as e:
            skipped.append({"pair_id": rec["pair_id"], "missing_feature": str(e)})
            continue
        if "log_eta" in rec:
            new["T"], new["log_eta"] = rec["T"], rec["log_eta"]
        if "mp" in rec:
            new["mp"] = rec["mp"]
        out.append(new)
    return out, skipped
