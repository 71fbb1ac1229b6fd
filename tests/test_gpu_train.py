"""GPU (B200): backward kernels and the optimiser step against the oracle's autograd (fp64) restatement of the
reference's training step (train_viscosity.py:227-230,328-338)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GRAD_RTOL = 2e-4  # fp32 kernels vs fp64 autograd, relative to the largest entry of each variable's gradient


def _setup(kind, n_pairs, seed, trained_like=True):
    from ionic_mpnn_b200 import graph, synth
    from ionic_mpnn_b200.model import MPNNModel
    from oracle import ref_inputs, ref_model

    label = "log_eta" if kind == "viscosity" else "mp"
    recs = synth.make_records(n_pairs, seed=seed, label=label)
    spec = ref_model.make_spec(kind)
    params = ref_model.init_params(spec, seed=seed + 1, trained_like=trained_like)
    y = np.array([r[label] for r in recs], np.float64)
    if kind != "viscosity":
        y = (y - y.mean()) / y.std()  # the reference z-scores the melting points outside the model
        for r, v in zip(recs, y):
            r[label] = float(v)
    x = ref_inputs.build_inputs(recs, with_temperature=kind == "viscosity")
    model = MPNNModel(spec, precision="fp32")
    model.set_weights(params)
    batch = graph.pack_records(recs, spec["bond_vocab_size"], label=label)
    return spec, params, x, y, model, batch


@pytest.mark.parametrize("kind", ["viscosity", "melting_point"])
def test_gradients_match_oracle_autograd(kind):
    from oracle import ref_model

    spec, params, x, y, model, batch = _setup(kind, 48, 3)
    loss_ref, grads_ref, out_ref = ref_model.loss_and_grads(spec, params, x, y)
    sse, out = model.loss_and_grads(batch)
    torch.cuda.synchronize()
    got = model.gradients()
    np.testing.assert_allclose(out.cpu().numpy(), out_ref[:, 0], rtol=2e-5, atol=2e-5)
    l2 = dict(ref_model.l2_terms(spec))
    mse_ref = float(((y - out_ref[:, 0]) ** 2).mean())
    assert abs(float(sse.item()) / len(y) - mse_ref) <= 1e-4 * max(1.0, mse_ref)
    worst = 0.0
    for k, gr in grads_ref.items():
        gr = gr - 2.0 * l2.get(k, 0.0) * np.asarray(params[k])  # the kernels leave the l2 term to imp_clip_adam
        scale = max(np.abs(gr).max(), 1e-8)
        err = np.abs(got[k] - gr).max() / scale
        worst = max(worst, err)
        assert err <= GRAD_RTOL, (k, err, scale)
    print(f"{kind}: worst relative gradient error {worst:.2e}; loss {loss_ref:.5f}")


def test_gradients_are_bit_reproducible():
    _, _, _, _, model, batch = _setup("viscosity", 300, 5)
    model.loss_and_grads(batch)
    torch.cuda.synchronize()
    a = model._train["grad"].clone()
    model.loss_and_grads(batch)
    torch.cuda.synchronize()
    assert torch.equal(a, model._train["grad"])


def test_train_steps_match_oracle_adam():
    """Three optimiser steps: parameters track the fp64 restatement of Keras' Adam(1e-3, clipnorm=1.0)."""
    from oracle import ref_model

    spec, params, x, y, model, batch = _setup("viscosity", 64, 7)
    p = {k: np.array(v, np.float64) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(w) for k, w in p.items()}
    losses_ref, losses = [], []
    clipped = set()
    for step in range(1, 4):
        loss_ref, grads, _, occ = ref_model.loss_and_grads(spec, p, x, y, occurrence_norms=True)
        norms = ref_model.adam_step(p, grads, m, v, step, occurrence_norm2=occ)  # Keras: per-occurrence norm for the Embeddings
        clipped |= {k for k, n in norms.items() if n > 1.0}
        losses_ref.append(loss_ref)
        losses.append(float(model.train_step(batch).item()))
    assert {"atom_emb", "bond_emb"} <= clipped, "the case must exercise the clip of both Embedding variables"
    np.testing.assert_allclose(losses, losses_ref, rtol=2e-4)
    got = model.get_weights()
    for k in p:
        err = np.abs(got[k] - p[k]).max()
        assert err <= 1e-4, (k, err)  # three steps of at most lr = 1e-3 each
    # the step really moved the weights and the loss went down on the training batch
    assert np.abs(got["head.bias"] - params["head.bias"]).max() > 1e-3
    assert losses[-1] < losses[0]


def test_embedding_occurrence_norms_match_oracle():
    """[Keras semantics] the clip norm of the two Embedding variables is taken over the per-occurrence gradient rows
    (oracle/ref_model.py:adam_step).  Atom ids and bond ids repeat in every batch, so it differs from the dense norm --
    here by more than 2x -- and the kernels (imp_sumsq, imp_bond_occurrence_norm2) must reproduce it."""
    from conftest import record_parity
    from oracle import ref_model

    for kind, n in (("viscosity", 48), ("viscosity", 333)):
        spec, params, x, y, model, batch = _setup(kind, n, 13)
        _, grads, _, occ = ref_model.loss_and_grads(spec, params, x, y, occurrence_norms=True)
        model.loss_and_grads(batch)
        torch.cuda.synchronize()
        tail = model._train["tail"].cpu().numpy().astype(np.float64)
        dense_atom = float((grads["atom_emb"] ** 2).sum())
        ea = abs(tail[2] - occ["atom_emb"]) / occ["atom_emb"]
        eb = abs(tail[3] - occ["bond_emb"]) / occ["bond_emb"]
        print(f"{n} pairs: occurrence norm^2 atom {tail[2]:.6g} (oracle {occ['atom_emb']:.6g}, dense {dense_atom:.6g}), "
              f"bond {tail[3]:.6g} (oracle {occ['bond_emb']:.6g}); rel err {ea:.2e} / {eb:.2e}")
        record_parity(f"train.occurrence_norm2.{n}_pairs", atom_rel_err=ea, bond_rel_err=eb)
        assert abs(dense_atom - occ["atom_emb"]) / occ["atom_emb"] > 0.5  # the two semantics really differ on this batch
        assert ea <= 1e-5 and eb <= 1e-3
    # run to run bit-identical (fixed reduction order)
    a = model._train["tail"].clone()
    model.loss_and_grads(batch)
    torch.cuda.synchronize()
    assert torch.equal(a, model._train["tail"])


def test_dense_clip_mode_is_the_old_semantics():
    """embedding_clip='dense' == oracle adam_step without occurrence norms (the deviation round 1 shipped)."""
    from oracle import ref_model

    spec, params, x, y, model, batch = _setup("viscosity", 64, 7)
    p = {k: np.array(v, np.float64) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(w) for k, w in p.items()}
    _, grads, _ = ref_model.loss_and_grads(spec, p, x, y)
    ref_model.adam_step(p, grads, m, v, 1)
    model.train_step(batch, embedding_clip="dense")
    got = model.get_weights()
    for k in p:
        assert np.abs(got[k] - p[k]).max() <= 2e-5, k


def test_half_batches_sum_to_the_full_batch_bucket():
    """SURVEY 4.5 on one GPU: two half batches with global_batch = the full size leave gradient buckets (and squared-error /
    occurrence-norm tails) whose SUM is the full batch's bucket -- the identity the one all-reduce of train_step relies on."""
    from ionic_mpnn_b200 import graph, synth

    spec, params, x, y, model, full = _setup("viscosity", 96, 21)
    recs = synth.make_records(96, seed=21, label="log_eta")
    halves = [graph.pack_records(recs[:50], 72, label="log_eta"), graph.pack_records(recs[50:], 72, label="log_eta")]
    model.loss_and_grads(full)
    torch.cuda.synchronize()
    want = model._train["grad"].clone()
    acc = torch.zeros_like(want)
    for hb in halves:
        model.loss_and_grads(hb, global_batch=96)
        torch.cuda.synchronize()
        acc += model._train["grad"]
    n = model.flat.numel()
    scale = want[:n].abs().max()
    assert float((acc[:n] - want[:n]).abs().max() / scale) <= 2e-6
    # tail: sse and the two occurrence norms add up as well
    for i in (0, 2, 3):
        assert abs(float(acc[n + i]) - float(want[n + i])) <= 2e-5 * abs(float(want[n + i])), i
    # "sum" mode (what train_step uses across ranks): sum-gradients / pair count == mean gradients
    model.loss_and_grads(full, global_batch="sum")
    torch.cuda.synchronize()
    assert float((model._train["grad"][:n] / 96.0 - want[:n]).abs().max() / scale) <= 2e-6


def test_asymmetric_batches_are_refused_for_training():
    """The message backward runs over the forward CSR, which is the true transpose only for symmetric live entries."""
    from ionic_mpnn_b200 import _lib, graph, synth
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    recs = synth.make_records(8, seed=2, label="log_eta")
    cat = graph.FlatIons.from_ion_dicts([r["cation"] for r in recs])
    an = graph.FlatIons.from_ion_dicts([r["anion"] for r in recs])
    keep = np.arange(len(cat.edge_src)) % 2 == 0   # featurize emits (a,b),(b,a) pairs: keep one direction only
    ep = np.zeros_like(cat.edge_ptr)
    ep[1:] = np.cumsum([keep[cat.edge_ptr[i]:cat.edge_ptr[i + 1]].sum() for i in range(cat.n_ions)])
    one_way = graph.FlatIons(cat.atom_ptr, cat.atom_ids, ep.astype(np.int32), np.ascontiguousarray(cat.edge_src[keep]),
                             np.ascontiguousarray(cat.edge_dst[keep]), np.ascontiguousarray(cat.bond_ids[keep]))
    y = np.array([r["log_eta"] for r in recs], np.float32)
    T = np.array([r["T"] for r in recs], np.float32)
    bad = graph.pack_flat(one_way, an, 72, double_edges=False, temperature=T, target=y)
    model = MPNNModel(make_spec("viscosity"), precision="fp32")
    with pytest.raises(_lib.ImpError, match="symmetric"):
        model.loss_and_grads(bad)
    good = graph.pack_flat(cat, an, 72, double_edges=False, temperature=T, target=y)  # both directions present: accepted
    model.loss_and_grads(good)
    assert good.symmetric is True and bad.symmetric is False


def test_training_then_fused_inference_uses_new_weights():
    """train_step invalidates the packed tensor-core weights; the fused forward picks the update up."""
    spec, params, x, y, model, batch = _setup("viscosity", 128, 9)
    from ionic_mpnn_b200.model import MPNNModel

    fused = MPNNModel(spec, precision="fp16")
    fused.set_weights(params)
    before = fused.forward_packed(batch).cpu().numpy()
    for _ in range(5):
        model.train_step(batch)
    fused.set_weights(model.get_weights())
    after = fused.forward_packed(batch).cpu().numpy()
    ref_after = model.forward_packed(batch).cpu().numpy()
    assert np.abs(after - before).max() > 1e-3
    assert np.abs(after - ref_after).max() <= 2e-2 * (np.abs(ref_after).max() + 1.0)


def test_fit_loop_early_stopping_and_best_weight_restore():
    """fit(): shuffled mini-batches of 32, sample-weighted epoch loss, val_loss per epoch, EarlyStopping(patience)
    with restore_best_weights (train_viscosity.py:328-338)."""
    from ionic_mpnn_b200 import synth
    from ionic_mpnn_b200.model import MPNNModel
    from oracle import ref_model

    spec = ref_model.make_spec("viscosity")
    train = synth.make_records(100, seed=11, label="log_eta")   # 3 batches of 32 + one of 4
    val = synth.make_records(40, seed=12, label="log_eta")      # labels are noise: val_loss stops improving early
    model = MPNNModel(spec, precision="fp32", seed=3)
    v0 = model.evaluate(val)
    hist = model.fit(train, validation_data=val, epochs=12, batch_size=32, patience=3, seed=1)
    assert len(hist["loss"]) == len(hist["val_loss"]) <= 12
    assert hist["loss"][-1] < hist["loss"][0]                    # the training loss goes down
    assert min(hist["val_loss"]) < v0
    stopped_early = len(hist["val_loss"]) < 12
    if stopped_early:                                            # stopped `patience` epochs after the best one
        assert len(hist["val_loss"]) - 1 - int(np.argmin(hist["val_loss"])) == 3
    # restore_best_weights: the model now holds the weights of the best epoch
    assert abs(model.evaluate(val) - min(hist["val_loss"])) <= 1e-5 * max(1.0, min(hist["val_loss"]))
    # determinism: same seed, same history
    again = MPNNModel(spec, precision="fp32", seed=3)
    h2 = again.fit(train, validation_data=val, epochs=len(hist["loss"]), batch_size=32, patience=3, seed=1)
    assert h2["loss"] == hist["loss"] and h2["val_loss"] == hist["val_loss"]


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """save_weights / load_weights (weights + Adam state): two steps, checkpoint, two more steps == four steps straight."""
    from ionic_mpnn_b200 import graph, synth
    from ionic_mpnn_b200.viscosity import build_model

    recs = synth.make_records(96, seed=3, label="log_eta")
    batch = graph.pack_records(recs, 72, label="log_eta")
    a = build_model(124, 72, seed=4)
    for _ in range(4):
        a.train_step(batch)
    b = build_model(124, 72, seed=4)
    for _ in range(2):
        b.train_step(batch)
    path = str(tmp_path / "ckpt.npz")
    b.save_weights(path)
    c = build_model(124, 72, seed=99)
    c.load_weights(path)
    for _ in range(2):
        c.train_step(batch)
    wa, wc = a.get_weights(), c.get_weights()
    for k in wa:
        assert np.array_equal(wa[k], wc[k]), k
    with pytest.raises(ValueError):
        build_model(124, 72, num_steps=3).load_weights(path)
