"""GPU (B200): the 3xTF32 tensor-core GatedUpdate forward (imp_gated_update_tc32, csrc/fwd_tc32.cu) against the fp32 SIMT kernel."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(n_atoms, n_cat, seed, gates=True):
    from ionic_mpnn_b200 import _lib

    rng = np.random.default_rng(seed)
    d = 32
    dev = "cuda"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)  # noqa: E731
    h, agg = t(rng.normal(size=(n_atoms, d))), t(rng.normal(size=(n_atoms, d)) * 2.0)
    ws, keep = [], []
    for _ in range(2):
        arrs = [t(rng.normal(size=(2 * d, d)) * 0.2), t(rng.normal(size=d) * 0.1), t(rng.normal(size=(2 * d, d)) * 0.2),
                t(rng.normal(size=d) * 0.1), t(rng.normal(size=(2 * d, d)) * 0.2), t(rng.normal(size=d) * 0.1),
                t(rng.uniform(0.5, 1.5, size=d)), t(rng.normal(size=d) * 0.1)]
        keep.append(arrs)
        ws.append(_lib.GruWeights(*[a.data_ptr() for a in arrs]))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for name in ("imp_gated_update_train", "imp_gated_update_tc32"):
        out, z, r, ht = (torch.full((n_atoms, d), 7.0, device=dev) for _ in range(4))
        if gates:
            _lib.call(name, h.data_ptr(), agg.data_ptr(), n_atoms, n_cat, d, C.byref(ws[0]), C.byref(ws[1]), C.c_float(1e-3),
                      out.data_ptr(), z.data_ptr(), r.data_ptr(), ht.data_ptr(), st)
        elif name == "imp_gated_update_tc32":
            _lib.call(name, h.data_ptr(), agg.data_ptr(), n_atoms, n_cat, d, C.byref(ws[0]), C.byref(ws[1]), C.c_float(1e-3),
                      out.data_ptr(), None, None, None, st)
        else:
            _lib.call("imp_gated_update", h.data_ptr(), agg.data_ptr(), n_atoms, n_cat, d, C.byref(ws[0]), C.byref(ws[1]),
                      C.c_float(1e-3), out.data_ptr(), st)
        torch.cuda.synchronize()
        res[name] = [x.cpu().numpy().astype(np.float64) for x in ((out, z, r, ht) if gates else (out,))]
    return res


@pytest.mark.parametrize("n_atoms,n_cat,seed", [(128, 64, 1), (1000, 517, 2), (40000, 21000, 3), (300, 0, 4), (300, 300, 5), (1, 1, 6)])
def test_tc32_forward_matches_the_fp32_kernel(n_atoms, n_cat, seed):
    res = _run(n_atoms, n_cat, seed)
    for nm, a, b in zip(("h_out", "z", "r", "ht"), res["imp_gated_update_train"], res["imp_gated_update_tc32"]):
        err = np.abs(a - b).max() / max(np.abs(a).max(), 1.0)
        print(nm, f"{err:.2e}")
        assert np.isfinite(b).all(), nm
        assert err <= 5e-6, (nm, err)  # fp32-class: summation order and the 3xTF32 split (~2^-21 per product; tanh of r*h adds r's error)


def test_tc32_plain_forward_and_reproducibility():
    a = _run(5000, 2600, 9, gates=False)
    err = np.abs(a["imp_gated_update_train"][0] - a["imp_gated_update_tc32"][0]).max()
    assert err <= 1e-5, err
    b = _run(5000, 2600, 9, gates=False)
    assert np.array_equal(a["imp_gated_update_tc32"][0], b["imp_gated_update_tc32"][0])
