"""CPU: host-side logic of the training path that needs no GPU -- the symmetry check of the live entry set (the message
backward runs over the forward CSR) and the oracle's restatement of Keras' clip semantics for Embedding variables."""
import numpy as np

from ionic_mpnn_b200 import graph, synth
from ionic_mpnn_b200.train import entries_are_symmetric
from oracle import ref_inputs, ref_model


def test_doubled_batches_are_symmetric_and_one_way_batches_are_not():
    recs = synth.make_records(20, seed=5)
    b = graph.pack_records(recs, 72)
    assert b.symmetric is True and entries_are_symmetric(b)
    cat = graph.FlatIons.from_ion_dicts([r["cation"] for r in recs])
    an = graph.FlatIons.from_ion_dicts([r["anion"] for r in recs])
    both = graph.pack_flat(cat, an, 72, double_edges=False)
    assert both.symmetric is None and entries_are_symmetric(both)          # featurize emits both directions
    keep = np.arange(len(cat.edge_src)) % 2 == 0
    ep = np.zeros_like(cat.edge_ptr)
    ep[1:] = np.cumsum([keep[cat.edge_ptr[i]:cat.edge_ptr[i + 1]].sum() for i in range(cat.n_ions)])
    one = graph.FlatIons(cat.atom_ptr, cat.atom_ids, ep.astype(np.int32), np.ascontiguousarray(cat.edge_src[keep]),
                         np.ascontiguousarray(cat.edge_dst[keep]), np.ascontiguousarray(cat.bond_ids[keep]))
    assert not entries_are_symmetric(graph.pack_flat(one, an, 72, double_edges=False))
    # truncation (max_edges) of a doubled batch can cut a mirror off: the flag is left to the check
    t = graph.pack_records(recs, 72, max_edges=5)
    assert t.symmetric is None


def test_occurrence_norm_differs_from_dense_norm_and_matches_a_manual_sum():
    """[Keras semantics] clip norm of an Embedding variable = norm over the un-deduplicated per-occurrence rows.  The
    oracle's taps are checked against a direct computation: the dense gradient is the scatter-add of the occurrence rows."""
    import torch

    recs = synth.make_records(6, seed=3, label="log_eta")
    spec = ref_model.make_spec("viscosity", atom_dim=8, bond_dim=4, fp_size=8, mixing_size=6, num_steps=2)
    params = ref_model.init_params(spec, seed=2, trained_like=True)
    x = ref_inputs.build_inputs(recs)
    y = np.array([r["log_eta"] for r in recs])
    p = ref_model.to_torch(params, torch.float64, requires_grad=True)
    taps = {}
    loss, _ = ref_model.loss_fn(spec, p, x, y, taps=taps)
    loss.backward()
    for var, key, ids in (("atom_emb", "atom_rows", "atom"), ("bond_emb", "bond_rows", "bond")):
        dense = torch.zeros_like(p[var])
        occ2 = 0.0
        for t in ("cat", "an"):
            g = taps[f"{t}_{key}"].grad
            idx = torch.as_tensor(np.asarray(x[f"{t}_{ids}"]), dtype=torch.long).reshape(-1)
            dense.index_add_(0, idx, g.reshape(-1, g.shape[-1]))
            occ2 += float((g ** 2).sum())
        assert torch.allclose(dense, p[var].grad, rtol=1e-10, atol=1e-14)
        assert abs(occ2 - float((p[var].grad ** 2).sum())) > 1e-3 * occ2  # ids repeat: the two norms differ
    _, grads, _, occ = ref_model.loss_and_grads(spec, params, x, y, occurrence_norms=True)
    # adam_step uses the occurrence norm where given and reports it
    pp = {k: np.array(v) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in pp.items()}
    v = {k: np.zeros_like(w) for k, w in pp.items()}
    norms = ref_model.adam_step(pp, grads, m, v, 1, occurrence_norm2=occ)
    assert abs(norms["atom_emb"] - np.sqrt(occ["atom_emb"])) < 1e-12
    assert abs(norms["head.bias"] - np.sqrt((grads["head.bias"] ** 2).sum())) < 1e-12
