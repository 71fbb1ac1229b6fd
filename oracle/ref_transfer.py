"""CPU restatement of the reference's transfer-learning model and losses (TEST INFRASTRUCTURE ONLY; parity unpinned against
a real TensorFlow run, like the rest of oracle/ -- see oracle/__init__.py).

Follows train_melting_point_transfer.py:
  build_transfer_model :76-106   viscosity graph up to "mix_cat_an", then Dense(256, relu) -> BatchNormalization ->
                                 Dense(128, relu) -> Dropout(0.3) -> Dense(64, relu) -> Dense(1)
  losses / optimizer   :195-197  tf.keras.losses.Huber(delta=1.0), Adam(lr) without clipnorm
Keras layer semantics restated (keras 2.12): BatchNormalization on a rank-2 input = non-fused path, batch statistics with the
biased variance, moving averages updated as m * 0.99 + batch * 0.01, epsilon 1e-3; Dropout keeps with probability 1 - rate and
scales by 1 / (1 - rate); Huber = mean over the batch of 0.5 e^2 (|e| <= delta) / delta (|e| - 0.5 delta).
The dropout mask is an INPUT here (Keras draws it from its own generator; the product draws it from a counter-based hash,
``dropout_keep`` below restates that hash so that both sides use the same mask).
"""
import numpy as np
import torch

from . import ref_model

BN_MOMENTUM, BN_EPS = 0.99, 1e-3


def dropout_keep(seed, n, rate):
    """keep mask of ionic_mpnn_b200/csrc/transfer_head.cu:dropout_uniform (splitmix64 of seed + golden * (i + 1))."""
    i = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * i
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u >= np.float32(rate)


def head_forward(hp, mixed, training, keep_mask=None, rate=0.3):
    """-> (prediction (B,1), new moving mean, new moving variance)."""
    a1 = torch.relu(mixed @ hp["mp_dense_1.kernel"] + hp["mp_dense_1.bias"])
    if training:
        mean = a1.mean(0)
        var = ((a1 - mean) ** 2).mean(0)  # biased
        mm = hp["mp_bn_1.moving_mean"].detach() * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM)
        mv = hp["mp_bn_1.moving_variance"].detach() * BN_MOMENTUM + var.detach() * (1 - BN_MOMENTUM)
    else:
        mean, var = hp["mp_bn_1.moving_mean"], hp["mp_bn_1.moving_variance"]
        mm, mv = mean, var
    bn = (a1 - mean) / torch.sqrt(var + BN_EPS) * hp["mp_bn_1.gamma"] + hp["mp_bn_1.beta"]
    a2 = torch.relu(bn @ hp["mp_dense_2.kernel"] + hp["mp_dense_2.bias"])
    if training and keep_mask is not None:
        a2 = a2 * torch.as_tensor(keep_mask.reshape(a2.shape), dtype=a2.dtype) / (1.0 - rate)
    a3 = torch.relu(a2 @ hp["mp_dense_3.kernel"] + hp["mp_dense_3.bias"])
    return a3 @ hp["melting_point.kernel"] + hp["melting_point.bias"], mm, mv


def huber(pred, y, delta=1.0):
    e = pred.reshape(-1) - y.reshape(-1)
    a = e.abs()
    return torch.where(a <= delta, 0.5 * e * e, delta * (a - 0.5 * delta)).mean()


def loss_and_grads(spec, base_params, head_params, x, y, trainable, keep_mask=None, rate=0.3, dtype=torch.float64):
    """Huber loss of the training-mode graph and d loss / d v for every variable name in ``trainable``."""
    p = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=(k in trainable)) for k, v in base_params.items()}
    hp = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=(k in trainable)) for k, v in head_params.items()}
    _, inter = ref_model.forward(spec, p, x, keep=True)
    pred, mm, mv = head_forward(hp, inter["mixed"], True, keep_mask, rate)
    loss = huber(pred, torch.as_tensor(np.asarray(y), dtype=dtype))
    loss.backward()
    grads = {k: v.grad.numpy() for k, v in {**p, **hp}.items() if v.requires_grad and v.grad is not None}
    return float(loss), grads, pred.detach().numpy(), mm.numpy(), mv.numpy()


def predict(spec, base_params, head_params, x, dtype=torch.float64):
    p = ref_model.to_torch(base_params, dtype)
    hp = ref_model.to_torch(head_params, dtype)
    with torch.no_grad():
        _, inter = ref_model.forward(spec, p, x, keep=True)
        return head_forward(hp, inter["mixed"], False)[0].numpy()


def adam_plain(w, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras 2.12 Adam._update_step without clipping (alpha = lr sqrt(1 - b2^t) / (1 - b1^t); v += (g^2 - v)(1 - b2))."""
    alpha = lr * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    m = m + (g - m) * (1 - beta1)
    v = v + (g * g - v) * (1 - beta2)
    return w - m * alpha / (np.sqrt(v) + eps), m, v
