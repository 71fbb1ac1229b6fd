"""GPU (B200): the tensor-core GatedUpdate backward (imp_gated_update_bwd_tc, csrc/bwd_tc.cu) against the fp32 SIMT kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(n_atoms, n_cat, seed):
    from ionic_mpnn_b200 import _lib

    rng = np.random.default_rng(seed)
    d = 32
    dev = "cuda"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)  # noqa: E731
    h, agg, gout = t(rng.normal(size=(n_atoms, d))), t(rng.normal(size=(n_atoms, d))), t(rng.normal(size=(n_atoms, d)) * 1e-2)
    ws, keep = [], []
    for _ in range(2):
        arrs = [t(rng.normal(size=(2 * d, d)) * 0.2), t(rng.normal(size=d) * 0.1), t(rng.normal(size=(2 * d, d)) * 0.2),
                t(rng.normal(size=d) * 0.1), t(rng.normal(size=(2 * d, d)) * 0.2), t(rng.normal(size=d) * 0.1),
                t(rng.uniform(0.5, 1.5, size=d)), t(rng.normal(size=d) * 0.1)]
        keep.append(arrs)
        ws.append(_lib.GruWeights(*[a.data_ptr() for a in arrs]))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out, z, r, ht = (torch.empty(n_atoms, d, device=dev) for _ in range(4))
    _lib.call("imp_gated_update_train", h.data_ptr(), agg.data_ptr(), n_atoms, n_cat, d, C.byref(ws[0]), C.byref(ws[1]), C.c_float(1e-3),
              out.data_ptr(), z.data_ptr(), r.data_ptr(), ht.data_ptr(), st)
    nws = _lib.load().imp_gated_update_bwd_workspace_floats(d)
    res = {}
    for name in ("imp_gated_update_bwd_stored", "imp_gated_update_bwd_tc"):
        dh, dagg = torch.full((n_atoms, d), 7.0, device=dev), torch.full((n_atoms, d), 7.0, device=dev)
        gc, ga = torch.zeros(3 * 2 * d * d + 5 * d, device=dev), torch.zeros(3 * 2 * d * d + 5 * d, device=dev)
        wsb = torch.zeros(nws, device=dev)
        _lib.call(name, h.data_ptr(), agg.data_ptr(), z.data_ptr(), r.data_ptr(), ht.data_ptr(), gout.data_ptr(), n_atoms, n_cat, d,
                  C.byref(ws[0]), C.byref(ws[1]), C.c_float(1e-3), dh.data_ptr(), dagg.data_ptr(), gc.data_ptr(), ga.data_ptr(),
                  wsb.data_ptr(), st)
        torch.cuda.synchronize()
        res[name] = [x.cpu().numpy().astype(np.float64) for x in (dh, dagg, gc, ga)]
    return res


@pytest.mark.parametrize("n_atoms,n_cat,seed", [(128, 64, 1), (1000, 517, 2), (40000, 21000, 3), (300, 0, 4), (300, 300, 5)])
def test_tensor_core_backward_matches_the_fp32_kernels(n_atoms, n_cat, seed):
    res = _run(n_atoms, n_cat, seed)
    ref, got = res["imp_gated_update_bwd_stored"], res["imp_gated_update_bwd_tc"]
    names = ("dh", "dagg", "grads_cat", "grads_an")
    for nm, a, b in zip(names, ref, got):
        scale = max(np.abs(a).max(), 1e-12)
        err = np.abs(a - b).max() / scale
        print(nm, f"{err:.2e}")
        assert np.isfinite(b).all(), nm
        assert err <= 2e-5, (nm, err)  # both are fp32-class; they differ by summation order and the 3xTF32 split (~2^-21)


def test_tensor_core_backward_is_bit_reproducible():
    a = _run(5000, 2600, 9)["imp_gated_update_bwd_tc"]
    b = _run(5000, 2600, 9)["imp_gated_update_bwd_tc"]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
