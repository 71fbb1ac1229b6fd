"""GPU (B200): the transfer-learning head and its two-stage training (train_melting_point_transfer.py) against
oracle/ref_transfer.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(n=48, seed=7):
    from ionic_mpnn_b200 import graph, synth
    from ionic_mpnn_b200.transfer import TransferModel, head_default_init
    from ionic_mpnn_b200.viscosity import build_model
    from oracle import ref_model

    recs = synth.make_records(n, seed=seed, n_min=4, n_max=14, label="mp")
    rng = np.random.default_rng(seed)
    y = rng.normal(0.0, 1.5, size=n).astype(np.float32)  # standardised labels with a few |e| > delta
    spec = ref_model.make_spec("viscosity")
    bp = ref_model.init_params(spec, seed=2, trained_like=True, bond_scale=4.0)
    hp = head_default_init(spec["mixing_size"], seed=3)
    hp["mp_bn_1.gamma"] = rng.uniform(0.5, 1.5, 256).astype(np.float32)
    hp["mp_bn_1.beta"] = rng.normal(0, 0.1, 256).astype(np.float32)
    hp["mp_bn_1.moving_mean"] = rng.normal(0, 0.1, 256).astype(np.float32)
    hp["mp_bn_1.moving_variance"] = rng.uniform(0.5, 1.5, 256).astype(np.float32)
    for k in hp:
        if k.endswith(".bias"):
            hp[k] = rng.normal(0, 0.05, hp[k].shape).astype(np.float32)
    base = build_model(124, 72, precision="fp32")
    base.set_weights({k: np.asarray(v, np.float32) for k, v in bp.items()})
    tm = TransferModel(base, seed=11)
    tm.set_head_weights(hp)
    x = None
    batch = graph.pack_records(recs, 72, label="mp")
    batch.target = y
    batch.to("cuda")
    return tm, spec, bp, hp, recs, x, y, batch


def _padded(recs):
    from oracle import ref_inputs

    x = ref_inputs.build_inputs(recs, with_temperature=False)
    x["temperature"] = np.full((len(recs), 1), 300.0, np.float32)  # the cut-off viscosity head still evaluates in the oracle
    return x


def test_transfer_inference_matches_the_oracle():
    from oracle import ref_transfer

    tm, spec, bp, hp, recs, _, y, batch = _setup()
    x = _padded(recs)
    want = ref_transfer.predict(spec, bp, hp, x).reshape(-1)
    got = tm.forward_packed(batch).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
    assert np.array_equal(tm.predict(recs).reshape(-1), got)  # records in, Keras-shaped (P, 1) out


@pytest.mark.parametrize("stage", [1, 2])
def test_transfer_gradients_match_fp64_autograd(stage):
    from ionic_mpnn_b200.transfer import UNFREEZE_KEYS
    from oracle import ref_transfer

    tm, spec, bp, hp, recs, _, y, batch = _setup()
    if stage == 2:
        tm.unfreeze(UNFREEZE_KEYS)
        names = tm.layer_names()
        unfrozen = sorted({names[v] for v in tm.trainable if v in tm.base.var_names})
        assert unfrozen == sorted(["cat_bmm_2", "cat_bmm_3", "an_bmm_2", "an_bmm_3", "gated_update_2", "gated_update_3",
                                   "gated_update_6", "gated_update_7"])
    x = _padded(recs)
    seed = 12345
    keep = ref_transfer.dropout_keep(seed, len(recs) * 128, 0.3)
    assert 0.6 < keep.mean() < 0.8
    loss_w, grads_w, pred_w, mm, mv = ref_transfer.loss_and_grads(spec, bp, hp, x, y, tm.trainable, keep)
    loss, pred = tm.loss_and_grads(batch, dropout_seed=seed)
    torch.cuda.synchronize()
    assert abs(float(loss.item()) / len(recs) - loss_w) <= 1e-5 * max(1.0, abs(loss_w))
    assert np.abs(pred.cpu().numpy() - pred_w.reshape(-1)).max() <= 2e-5 * max(1.0, np.abs(pred_w).max())
    gh = tm._opt["gh"].cpu().numpy()
    gb = tm.base.gradients() if stage == 2 else {}
    for k in sorted(tm.trainable):
        want = grads_w[k]
        if k in tm.off:
            got = gh[tm.off[k]: tm.off[k] + want.size].reshape(want.shape)
        else:
            got = gb[k]
        scale = max(np.abs(want).max(), 1e-12)
        assert np.abs(got - want).max() <= 2e-4 * scale, (k, np.abs(got - want).max() / scale)
    # the moving averages moved the way Keras moves them
    assert np.abs(tm.params["mp_bn_1.moving_mean"].cpu().numpy() - mm).max() <= 1e-5
    assert np.abs(tm.params["mp_bn_1.moving_variance"].cpu().numpy() - mv).max() <= 1e-5


def test_two_stage_training_freezes_by_layer_name_and_tracks_the_oracle():
    """Stage 1: only mp_* / melting_point variables change.  Stage 2 (fresh optimizer, UNFREEZE_KEYS): the last two message
    steps of both towers change too, everything else stays bit-identical.  Three Adam steps per stage follow the fp64
    restatement (same dropout masks)."""
    from ionic_mpnn_b200.transfer import UNFREEZE_KEYS
    from oracle import ref_transfer

    tm, spec, bp, hp, recs, _, y, batch = _setup(n=40, seed=9)
    x = _padded(recs)
    w_base = {k: np.asarray(v, np.float64) for k, v in bp.items()}
    w_head = {k: np.asarray(v, np.float64) for k, v in hp.items()}
    for stage, lr in ((1, 1e-3), (2, 1e-4)):
        if stage == 2:
            tm.unfreeze(UNFREEZE_KEYS)
        tm.compile()
        m = {k: 0.0 for k in tm.trainable}
        v = {k: 0.0 for k in tm.trainable}
        before_b, before_h = tm.base.get_weights(), tm.get_head_weights()
        for step in range(1, 4):
            seed = 1000 * stage + step
            keep = ref_transfer.dropout_keep(seed, len(recs) * 128, 0.3)
            loss_w, grads, _, mm, mv = ref_transfer.loss_and_grads(spec, w_base, w_head, x, y, tm.trainable, keep)
            for k in tm.trainable:
                tgt = w_head if k in w_head else w_base
                tgt[k], m[k], v[k] = ref_transfer.adam_plain(tgt[k], grads[k], m[k], v[k], step, lr)
            w_head["mp_bn_1.moving_mean"], w_head["mp_bn_1.moving_variance"] = mm, mv
            loss = tm.train_step(batch, lr=lr, dropout_seed=seed)
            assert abs(float(loss.item()) - loss_w) <= 2e-5 * max(1.0, abs(loss_w)), (stage, step)
        after_b, after_h = tm.base.get_weights(), tm.get_head_weights()
        for k in after_b:
            if k in tm.trainable:
                assert not np.array_equal(after_b[k], before_b[k]), k
                assert np.abs(after_b[k] - w_base[k]).max() <= 2e-4 * max(1.0, np.abs(w_base[k]).max()), (stage, k)
            else:
                assert np.array_equal(after_b[k], before_b[k]), (stage, k)
        for k in after_h:
            assert np.abs(after_h[k] - w_head[k]).max() <= 2e-4 * max(1.0, np.abs(w_head[k]).max()), (stage, k)


def test_transfer_archive_round_trip_and_fused_inference(tmp_path):
    """The head runs on top of whatever forward path the base uses: fp16 fused base within the 16-bit tolerance of fp32."""
    from ionic_mpnn_b200.transfer import TransferModel
    from ionic_mpnn_b200.viscosity import build_model

    tm, spec, bp, hp, recs, _, y, batch = _setup(n=300, seed=4)
    ref = tm.forward_packed(batch).cpu().numpy()
    fast = build_model(124, 72, precision="fp16", fused=True)
    fast.set_weights(tm.base.get_weights())
    tm2 = TransferModel(fast)
    tm2.set_head_weights(tm.get_head_weights())
    got = tm2.forward_packed(batch).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max())
    # model.save / load_model of the transfer model (train_melting_point_transfer.py:243-251)
    path = str(tmp_path / "transfer.keras")
    tm.save_keras(path)
    tm3 = TransferModel(build_model(124, 72, precision="fp32", seed=99), seed=98).load_keras(path)
    assert np.array_equal(tm3.forward_packed(batch).cpu().numpy(), ref)
