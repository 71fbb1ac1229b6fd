import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a machine without CUDA or without the built library, so that a
    plain `pytest tests` works everywhere; on a GPU box nothing is skipped (there is no CPU fallback to hide behind)."""
    import torch

    from ionic_mpnn_b200 import _lib

    if torch.cuda.is_available() and os.path.exists(_lib.LIB_PATH):
        return
    why = "needs CUDA" if not torch.cuda.is_available() else "libimp_b200.so is not built"
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# Measured parity errors of a `-m gpu` run, collected by tests through `record_parity` and written at session end to
# gpurun_out/parity_run.json (copied to profiles/rNN_parity.json by hand after a GPU run).
_PARITY = {}


def record_parity(key, **values):
    _PARITY.setdefault(key, {}).update({k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in values.items()})


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    try:
        import torch

        dev = torch.cuda.get_device_name(0) if torch.cuda.is_available() else None
    except Exception:  # pragma: no cover
        dev = None
    with open(os.path.join(out, "parity_run.json"), "w") as f:
        json.dump({"device": dev, "metric": "max |got - want| / max(|want|, 1) unless a key says otherwise",
                   "exitstatus": int(exitstatus), "cases": _PARITY}, f, indent=1, sort_keys=True)


def load_golden(name):
    """Returns (meta, x, inter, out, params) of one tests/golden/*.npz case (see make_golden.py)."""
    from oracle import ref_model

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    x = {k[2:]: z[k] for k in z.files if k.startswith("x.")}
    inter = {k[2:]: z[k] for k in z.files if k.startswith("i.")}
    if meta["weights_stored"]:
        params = {k[2:]: z[k].astype(np.float64) for k in z.files if k.startswith("w.")}
    else:
        params = ref_model.init_params(meta["spec"], **meta["init"])
    for r in meta["records"]:  # json turned the edge tuples into lists
        for ion in ("cation", "anion"):
            r[ion]["edge_indices"] = [tuple(e) for e in r[ion]["edge_indices"]]
    return meta, x, inter, z["out"], params


GOLDEN_F64 = ["visc_default_init", "visc_trained_like", "visc_small", "visc_batched_predict", "mp_small",
              "mp_default_dims"]


@pytest.fixture(scope="session")
def golden():
    return load_golden
