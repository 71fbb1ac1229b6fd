"""TEST INFRASTRUCTURE ONLY -- a numpy stand-in for the TensorFlow / Keras API surface
that goalheart/ionic-mpnn touches on its MPNN hot path.

Purpose: TensorFlow 2.12 (the reference's pinned runtime, environment.yml:10) is not
installed and cannot be installed here.  This shim lets ``tests/golden/make_golden.py``
execute the reference's OWN Python (``models/layers.py`` verbatim; ``build_model`` and the
padding helpers lifted from the training scripts) so that the golden vectors are produced
by the reference's wiring rather than by a second hand-written copy of it.

It is deliberately tiny and eager: a "tensor" is a numpy array, the Keras functional API
is a recorded DAG of ``Sym`` nodes that ``Model.predict`` evaluates batch by batch.
Library semantics encoded here ([Keras semantics] in SURVEY.md section 8c):

* ``Embedding``: uniform(-0.05, 0.05) init, plain row gather, ``mask_zero`` ignored when False.
* ``Dense``: glorot_uniform kernel ``(in, out)``, zero bias, ``y = x @ kernel + bias``.
* ``add_weight(initializer="glorot_uniform")`` for rank-3 ``(K, d, d)``:
  fan_in = shape[-2] * prod(shape[:-2]), fan_out = shape[-1] * prod(shape[:-2]).
* ``LayerNormalization``: last axis, epsilon 1e-3, biased variance, gamma 1, beta 0.
* ``Dropout``: identity at inference.
* ``softplus(x) = log1p(exp(x))`` (computed stably), ``sigmoid``, ``tanh``.
* ``Model.predict``: default batch size 32, concatenated along axis 0.

Not covered (never used for parity claims): training, autodiff, saving.
"""
from __future__ import annotations

import re
import sys
import types

import numpy as np

_STATE = {"float": np.float64, "rng": np.random.default_rng(0), "layers": [], "names": {}}


def configure(float_dtype=np.float64, seed=0):
    """Select the working float dtype (the reference hard-codes tf.float32; fp64 gives a
    tighter pin) and reset the initialiser RNG and the layer registry."""
    _STATE["float"] = np.dtype(float_dtype).type
    _STATE["rng"] = np.random.default_rng(seed)
    _STATE["layers"] = []
    _STATE["names"] = {}


def created_layers():
    return list(_STATE["layers"])


# ----------------------------------------------------------------------------- tf.* ops
class _DType:
    def __init__(self, name):
        self.name = name

    def np(self):
        if self.name == "float32":
            return _STATE["float"]  # working float dtype (fp32 or fp64)
        return {"int32": np.int32, "int64": np.int64, "bool": np.bool_}[self.name]


float32 = _DType("float32")
int32 = _DType("int32")
int64 = _DType("int64")


def _npdt(dt):
    return dt.np() if isinstance(dt, _DType) else dt


def shape(x):
    return np.asarray(np.shape(x), dtype=np.int64)


def range_(n, dtype=int32):
    return np.arange(int(n), dtype=_npdt(dtype))


def tile(x, multiples):
    return np.tile(x, [int(m) for m in multiples])


def stack(xs, axis=0):
    return np.stack(xs, axis=axis)


def reshape(x, shp):
    return np.reshape(x, [int(s) for s in shp])


def boolean_mask(t, mask):
    return np.asarray(t)[np.asarray(mask, dtype=bool)]


def scatter_nd(indices, updates, shape):
    shp = tuple(int(s) for s in shape)
    out = np.zeros(shp, dtype=np.asarray(updates).dtype)
    idx = np.asarray(indices)
    if idx.shape[0]:
        np.add.at(out, tuple(idx[:, j] for j in range(idx.shape[1])), updates)
    return out


def gather(params, indices, batch_dims=0, axis=None):
    params = np.asarray(params)
    indices = np.asarray(indices)
    if batch_dims == 0:
        return np.take(params, indices, axis=0 if axis is None else axis)
    assert batch_dims == 1 and indices.ndim == 2
    b = np.arange(params.shape[0])[:, None]
    return params[b, indices]


def tensordot(a, b, axes):
    return np.tensordot(a, b, axes=axes)


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def matmul(a, b):
    return np.matmul(a, b)


def squeeze(x, axis=None):
    return np.squeeze(x, axis=axis)


def logical_and(a, b):
    return np.logical_and(a, b)


def cast(x, dtype):
    return np.asarray(x).astype(_npdt(dtype))


def concat(xs, axis):
    return np.concatenate(xs, axis=axis)


def reduce_sum(x, axis=None):
    return np.sum(x, axis=axis)


def clip_by_value(x, lo, hi):
    return np.clip(x, lo, hi)


def _softplus(x):
    return np.logaddexp(x, 0.0)  # == log1p(exp(x)), overflow-safe


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# ------------------------------------------------------------------- Keras functional API
class Sym:
    """A node of the recorded functional graph."""

    def __init__(self, fn, parents, name=None):
        self.fn, self.parents, self.name = fn, parents, name

    def __getitem__(self, key):
        return Sym(lambda x: x[key], [self])


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for y in x:
            out.extend(_flatten(y))
        return out
    return [x]


def _evaluate(node, feed, cache):
    if not isinstance(node, Sym):
        return node
    if id(node) in cache:
        return cache[id(node)]
    if node.fn is None:
        val = np.asarray(feed[node.name])
        if getattr(node, "dtype", None) is not None:
            val = val.astype(_npdt(node.dtype))
    else:
        val = node.fn(*[_evaluate(p, feed, cache) for p in node.parents])
    cache[id(node)] = val
    return val


def _auto_name(cls):
    base = re.sub(r"(?<!^)(?=[A-Z])", "_", cls.__name__).lower()
    k = _STATE["names"].get(base, 0)
    _STATE["names"][base] = k + 1
    return base if k == 0 else f"{base}_{k}"


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name or _auto_name(type(self))
        self.built = False
        self.weights = {}
        self.output = None  # last Sym produced (functional mode)
        _STATE["layers"].append(self)

    def add_weight(self, shape=None, initializer="glorot_uniform", name=None, **kw):
        shape = tuple(int(s) for s in shape)
        f = _STATE["float"]
        rng = _STATE["rng"]
        if initializer == "glorot_uniform":
            if len(shape) == 2:
                fan_in, fan_out = shape
            else:  # Keras _compute_fans for rank>2: receptive field = prod(shape[:-2])
                rf = int(np.prod(shape[:-2]))
                fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-lim, lim, size=shape).astype(f)
        elif initializer == "uniform":
            w = rng.uniform(-0.05, 0.05, size=shape).astype(f)
        elif initializer == "zeros":
            w = np.zeros(shape, f)
        elif initializer == "ones":
            w = np.ones(shape, f)
        else:
            raise NotImplementedError(initializer)
        self.weights[name] = w
        return _WeightRef(self, name)

    def build(self, input_shape):
        pass

    def call(self, inputs, **kw):
        raise NotImplementedError

    def get_config(self):
        return {"name": self.name}

    def _eager(self, inputs, **kw):
        if not self.built:
            self.build(None if isinstance(inputs, (list, tuple)) else np.shape(inputs))
            self.built = True
        return self.call(inputs, **kw)

    def __call__(self, inputs, **kw):
        flat = _flatten(inputs)
        if any(isinstance(x, Sym) for x in flat):
            is_list = isinstance(inputs, (list, tuple))

            def fn(*vals):
                return self._eager(list(vals) if is_list else vals[0], **kw)

            self.output = Sym(fn, flat, name=self.name)
            return self.output
        return self._eager(inputs, **kw)


class _WeightRef:
    """Stands for a tf.Variable: reads through to the owning layer so weights can be
    replaced after build (``layer.weights[name] = ...``)."""

    def __init__(self, layer, name):
        self.layer, self.wname = layer, name

    def __array__(self, dtype=None, copy=None):
        a = self.layer.weights[self.wname]
        return a if dtype is None else a.astype(dtype)

    @property
    def shape(self):
        return self.layer.weights[self.wname].shape


class Dense(Layer):
    def __init__(self, units, activation=None, kernel_regularizer=None, **kw):
        super().__init__(**kw)
        self.units, self.activation, self.kernel_regularizer = units, activation, kernel_regularizer

    def _eager(self, inputs, **kw):
        if not self.built:
            self.add_weight((np.shape(inputs)[-1], self.units), "glorot_uniform", "kernel")
            self.add_weight((self.units,), "zeros", "bias")
            self.built = True
        y = np.matmul(inputs, self.weights["kernel"]) + self.weights["bias"]
        if self.activation == "relu":
            y = np.maximum(y, 0)
        elif self.activation is not None:
            raise NotImplementedError(self.activation)
        return y


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, mask_zero=False, **kw):
        super().__init__(**kw)
        assert not mask_zero
        self.add_weight((input_dim, output_dim), "uniform", "embeddings")
        self.built = True

    def call(self, ids, **kw):
        return self.weights["embeddings"][np.asarray(ids)]


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.epsilon = epsilon

    def _eager(self, x, **kw):
        if not self.built:
            self.add_weight((np.shape(x)[-1],), "ones", "gamma")
            self.add_weight((np.shape(x)[-1],), "zeros", "beta")
            self.built = True
        mean = np.mean(x, axis=-1, keepdims=True)
        var = np.mean((x - mean) ** 2, axis=-1, keepdims=True)
        inv = 1.0 / np.sqrt(var + _STATE["float"](self.epsilon))
        return (x - mean) * inv * self.weights["gamma"] + self.weights["beta"]


class Dropout(Layer):
    def __init__(self, rate, **kw):
        super().__init__(**kw)
        self.rate = rate

    def call(self, x, training=None, **kw):
        assert not training
        return x


class Add(Layer):
    def call(self, xs, **kw):
        out = xs[0]
        for x in xs[1:]:
            out = out + x
        return out


def Input(shape=None, dtype=None, name=None):
    node = Sym(None, [], name=name)
    node.dtype = dtype  # float inputs are cast to the working float dtype at feed time
    return node


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs, self.name = list(inputs), outputs, name
        self.layers = created_layers()

    def compile(self, **kw):
        pass

    def get_layer(self, name):
        return next(l for l in self.layers if l.name == name)

    def run(self, feed, fetch):
        """Evaluate arbitrary Sym nodes for one feed dict (no batching)."""
        cache = {}
        return [_evaluate(n, feed, cache) for n in fetch]

    def predict(self, x, batch_size=32, verbose=0):
        n = len(next(iter(x.values())))
        outs = []
        for s in range(0, n, batch_size):
            feed = {k: np.asarray(v)[s : s + batch_size] for k, v in x.items()}
            outs.append(self.run(feed, [self.outputs])[0])
        return np.concatenate(outs, axis=0)

    __call__ = predict


def register_keras_serializable(*a, **kw):
    return lambda cls: cls


def install():
    """Register the shim as ``tensorflow`` (+ the keras submodules the reference imports)."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.int64 = float32, int32, int64
    for f in (shape, tile, stack, reshape, boolean_mask, scatter_nd, gather, tensordot, expand_dims,
              matmul, squeeze, logical_and, cast, concat, reduce_sum, clip_by_value):
        setattr(tf, f.__name__, f)
    tf.range = range_
    nn = types.ModuleType("tensorflow.nn")
    nn.softplus, nn.sigmoid, nn.tanh = _softplus, _sigmoid, np.tanh
    tf.nn = nn
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    for c in (Layer, Dense, Embedding, LayerNormalization, Dropout, Add):
        setattr(layers, c.__name__, c)
    layers.Input = Input
    saving = types.ModuleType("tensorflow.keras.saving")
    saving.register_keras_serializable = register_keras_serializable
    models = types.ModuleType("tensorflow.keras.models")
    models.Model = Model
    callbacks = types.ModuleType("tensorflow.keras.callbacks")
    callbacks.EarlyStopping = type("EarlyStopping", (), {"__init__": lambda self, **kw: None})
    callbacks.Callback = type("Callback", (), {})
    regularizers = types.ModuleType("tensorflow.keras.regularizers")
    regularizers.l2 = lambda v: ("l2", v)
    optimizers = types.ModuleType("tensorflow.keras.optimizers")
    optimizers.Adam = lambda *a, **kw: ("adam", a, kw)
    keras.layers, keras.saving, keras.models = layers, saving, models
    keras.callbacks, keras.regularizers, keras.optimizers = callbacks, regularizers, optimizers
    tf.keras = keras
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.keras": keras,
            "tensorflow.keras.layers": layers, "tensorflow.keras.saving": saving,
            "tensorflow.keras.models": models, "tensorflow.keras.callbacks": callbacks,
            "tensorflow.keras.regularizers": regularizers, "tensorflow.keras.optimizers": optimizers}
    sys.modules.update(mods)
    return tf
