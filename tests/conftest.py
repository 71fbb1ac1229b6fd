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
