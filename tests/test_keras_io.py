"""CPU: the dependency-free HDF5 reader / writer and the ``.keras`` archive reader."""
import json
import os
import zipfile

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from ionic_mpnn_b200 import hdf5_min, keras_io


def test_hdf5_round_trip_groups_dtypes_shapes():
    rng = np.random.default_rng(0)
    data = {"a/b/vars/0": rng.normal(size=(3, 5)).astype(np.float32), "a/b/vars/1": rng.normal(size=(7,)).astype(np.float64),
            "a/c": np.arange(12, dtype=np.int32).reshape(2, 3, 2), "scalar": np.float32(3.5),
            "empty": np.zeros((0, 4), np.float32), "z/u8": np.arange(5, dtype=np.uint8), "z/i64": np.array([-1, 2 ** 40], np.int64)}
    img = hdf5_min.write_datasets(data)
    assert img[:8] == b"\x89HDF\r\n\x1a\n"
    back = hdf5_min.read_datasets(img)
    assert set(back) == set(data)
    for k, v in data.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape, k
        assert np.array_equal(back[k], v), k
    many = {f"g/{i:03d}": np.full((2,), i, np.float32) for i in range(300)}  # one group with 300 links
    assert all(np.array_equal(v, many[k]) for k, v in hdf5_min.read_datasets(hdf5_min.write_datasets(many)).items())


def test_hdf5_reader_rejects_garbage_and_unsupported_features():
    with pytest.raises(hdf5_min.H5Error):
        hdf5_min.read_datasets(b"not an hdf5 file" * 10)
    img = bytearray(hdf5_min.write_datasets({"x": np.ones(3, np.float32)}))
    img[8] = 9  # superblock version
    with pytest.raises(hdf5_min.H5Error, match="superblock"):
        hdf5_min.read_datasets(bytes(img))


@pytest.mark.parametrize("fixture,golden", [("visc_small.keras", "visc_small"), ("visc_small_by_name.keras", "visc_small"),
                                            ("mp_small.keras", "mp_small")])
def test_archives_of_the_reference_graph_load_by_structure(fixture, golden):
    """tests/golden/*.keras serialise the graph that the reference's own build_model records (make_keras_fixture.py);
    the reader must recover exactly the weights that were injected by structure when the golden vectors were made."""
    config, data = keras_io.read_keras(os.path.join(GOLDEN_DIR, fixture))
    kind, params, extra = keras_io.params_from_keras(config, data)
    z = np.load(os.path.join(GOLDEN_DIR, golden + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    want = {k[2:]: z[k].astype(np.float32) for k in z.files if k.startswith("w.")}
    assert kind == meta["kind"] and not extra
    assert set(params) == set(want)
    for k in want:
        assert np.array_equal(params[k], want[k]), k
    assert keras_io.spec_from_params(kind, params) == meta["spec"]
    names = [l["name"] for l in config["config"]["layers"]]
    assert "gated_update_5" in names and "cat_reduce_0" in names  # auto-named and explicitly named layers, as in Keras


@pytest.mark.parametrize("kind", ["viscosity", "melting_point"])
@pytest.mark.parametrize("key_style", ["class_counter", "layer_name"])
def test_export_then_read_is_bit_exact(tmp_path, kind, key_style):
    from ionic_mpnn_b200.model import keras_default_init, make_spec, param_shapes

    spec = make_spec(kind, atom_dim=8, bond_dim=4, fp_size=8, mixing_size=6, num_steps=2)
    params = keras_default_init(spec, seed=5)
    rng = np.random.default_rng(1)
    params = {k: (v + rng.normal(size=v.shape).astype(np.float32)) for k, v in params.items()}
    path = str(tmp_path / "m.keras")
    keras_io.export_keras(path, spec, params, key_style=key_style)
    with zipfile.ZipFile(path) as z:
        assert {"config.json", "metadata.json", "model.weights.h5"} <= set(z.namelist())
    config, data = keras_io.read_keras(path)
    k2, back, _ = keras_io.params_from_keras(config, data)
    assert k2 == kind and set(back) == set(param_shapes(spec))
    for k, v in params.items():
        assert np.array_equal(back[k], v), k
    assert keras_io.spec_from_params(k2, back) == spec


def test_reader_reports_a_missing_layer_by_name(tmp_path):
    from ionic_mpnn_b200.model import keras_default_init, make_spec

    spec = make_spec("viscosity", atom_dim=8, bond_dim=4, fp_size=8, mixing_size=6, num_steps=2)
    path = str(tmp_path / "m.keras")
    keras_io.export_keras(path, spec, keras_default_init(spec))
    config, data = keras_io.read_keras(path)
    data = {k: v for k, v in data.items() if "bond_matrix_message_1" not in k}
    with pytest.raises(ValueError, match="cat_bmm_1"):
        keras_io.params_from_keras(config, data)


def test_transfer_archive_keeps_the_named_head_layers(tmp_path):
    """build_transfer_model's archive (train_melting_point_transfer.py:76-106): the base is cut at mix_cat_an (no Dense(3)
    head), the mp_* / melting_point layers carry their own names and come back as ``extra`` in Keras variable order."""
    from ionic_mpnn_b200.model import keras_default_init, make_spec
    from ionic_mpnn_b200.transfer import head_default_init

    spec = make_spec("viscosity", atom_dim=8, bond_dim=4, fp_size=8, mixing_size=6, num_steps=2)
    base = {k: v for k, v in keras_default_init(spec, seed=2).items() if not k.startswith("head")}
    head = head_default_init(spec["mixing_size"], seed=3)
    head["mp_bn_1.moving_mean"] = np.linspace(-1, 1, 256).astype(np.float32)
    path = str(tmp_path / "t.keras")
    keras_io.export_keras(path, spec, {**base, **head}, transfer_head=True)
    config, data = keras_io.read_keras(path)
    kind, params, extra = keras_io.params_from_keras(config, data)
    assert kind == "transfer" and set(params) == set(base)
    assert all(np.array_equal(params[k], base[k]) for k in base)
    assert set(extra) == {"mp_dense_1", "mp_bn_1", "mp_dense_2", "mp_dense_3", "melting_point"}
    assert np.array_equal(extra["mp_bn_1"][2], head["mp_bn_1.moving_mean"]) and len(extra["mp_bn_1"]) == 4
    assert np.array_equal(extra["melting_point"][0], head["melting_point.kernel"])
