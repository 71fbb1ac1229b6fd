"""Golden fixture of the vocabulary / id-record stage, generated HERE by running the reference's own code
(/root/reference/src/build_vocab.py, src/dataset.py) on synthetic feature-tuple graphs.

    python tests/golden/make_featurize_fixture.py        (needs /root/reference; writes tests/golden/featurize_small.json)
"""
import json
import os
import pickle
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SYMS = ["C", "N", "O", "H"]
HYB = ["SP2", "SP3", "S"]
BT = ["SINGLE", "DOUBLE", "TRIPLE", "AROMATIC"]


def graph(rng, n):
    atoms = [(rng.choice(SYMS), rng.choice([0, 0, 0, 1]), rng.randrange(0, 3), rng.randrange(0, 2), rng.choice(HYB)) for _ in range(n)]
    bf, ei = [], []
    for i in range(1, n):
        j = rng.randrange(0, i)
        f = (rng.choice(BT), rng.random() < 0.4, rng.random() < 0.3)
        ei += [(j, i), (i, j)]
        bf += [f, f]
    return {"smiles": "synthetic", "atom_features": atoms, "bond_features": bf, "edge_indices": ei, "num_atoms": n}


def records(rng, n, label):
    out = []
    for i in range(n):
        r = {"pair_id": f"{label}{i}", "cation_graph": graph(rng, rng.randrange(1, 12)), "anion_graph": graph(rng, rng.randrange(1, 9))}
        if label == "v":
            r["T"], r["log_eta"] = 273.15 + rng.random() * 100, rng.gauss(2, 1)
        else:
            r["mp"] = rng.gauss(330, 60)
        out.append(r)
    return out


def main():
    rng = random.Random(7)
    vis, mp = records(rng, 30, "v"), records(rng, 20, "m")
    extra = records(rng, 6, "v")  # processed against the vocabulary of vis + mp: two of these hold unseen features
    extra[1]["cation_graph"]["atom_features"][0] = ("Xe", 0, 0, 0, "SP3D2")
    extra[4]["anion_graph"]["bond_features"][:2] = [("QUADRUPLE", False, False)] * 2 if extra[4]["anion_graph"]["bond_features"] else []
    if not extra[4]["anion_graph"]["bond_features"]:
        extra[4]["anion_graph"]["atom_features"][0] = ("Xe", 0, 0, 0, "SP3D2")
    sys.path.insert(0, "/root/reference/src")
    import build_vocab as ref_vocab  # noqa: E402
    import dataset as ref_dataset  # noqa: E402

    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        os.makedirs("data")
        for name, d in (("viscosity_graph_data.pkl", vis), ("mp_graph_data.pkl", mp), ("extra_graph_data.pkl", extra)):
            with open(os.path.join("data", name), "wb") as f:
                pickle.dump(d, f)
        vocab = ref_vocab.build_vocab_from_graph_data("data/viscosity_graph_data.pkl", "data/mp_graph_data.pkl")
        outs = {}
        for name in ("viscosity", "mp", "extra"):
            skipped = ref_dataset.process_dataset(f"data/{name}_graph_data.pkl", "data/vocab.pkl", f"data/{name}_id_data.pkl")
            with open(f"data/{name}_id_data.pkl", "rb") as f:
                outs[name] = {"records": pickle.load(f), "skipped": skipped}
        os.chdir(cwd)
    fix = {"inputs": {"viscosity": vis, "mp": mp, "extra": extra},
           "vocab": {"atom_vocab": [[list(k), v] for k, v in vocab["atom_vocab"].items()],
                     "bond_vocab": [[list(k), v] for k, v in vocab["bond_vocab"].items()],
                     "atom_vocab_size": vocab["atom_vocab_size"], "bond_vocab_size": vocab["bond_vocab_size"]},
           "outputs": outs}
    with open(os.path.join(HERE, "featurize_small.json"), "w") as f:
        json.dump(fix, f)
    print("wrote featurize_small.json:", vocab["atom_vocab_size"], "atom /", vocab["bond_vocab_size"], "bond features;",
          {k: (len(v["records"]), len(v["skipped"])) for k, v in outs.items()})


if __name__ == "__main__":
    main()
