"""GPU (B200): the tcgen05 / TMEM plumbing in isolation (imp_tc_selftest) against a CPU product."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


def _tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def run_selftest(N, K, kind, swap=0, seed=0):
    from ionic_mpnn_b200 import _lib

    rng = np.random.default_rng(seed)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dD = torch.zeros(128, N, dtype=torch.float32, device="cuda")
    _lib.call("imp_tc_selftest", dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), N, K, kind, swap,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return A, B, dD.cpu().numpy()


@pytest.mark.parametrize("N,K", [(32, 32), (64, 64), (32, 256), (64, 16)])
def test_bf16_mma_matches_cpu(N, K):
    A, B, D = run_selftest(N, K, 0)
    want = _bf16_round(A).astype(np.float64) @ _bf16_round(B).astype(np.float64).T
    assert np.abs(D - want).max() <= 1e-3 * max(1.0, np.abs(want).max()), np.abs(D - want).max()


@pytest.mark.parametrize("N,K", [(32, 32), (64, 64), (32, 8)])
def test_tf32_mma_matches_cpu(N, K):
    A, B, D = run_selftest(N, K, 1)
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    # tf32 keeps 10 mantissa bits: error per product ~2^-11 relative (round or truncate, hardware's choice)
    assert np.abs(D - exact).max() <= 4e-3 * np.sqrt(K) * 4.0, np.abs(D - exact).max()
    # and an exactly representable input must give the exact answer
    A2, B2 = _tf32_trunc(A), _tf32_trunc(B)
    from ionic_mpnn_b200 import _lib

    dA, dB = torch.from_numpy(A2).cuda(), torch.from_numpy(B2).cuda()
    dD = torch.zeros(128, N, dtype=torch.float32, device="cuda")
    _lib.call("imp_tc_selftest", dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), N, K, 1, 0,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = A2.astype(np.float64) @ B2.astype(np.float64).T
    assert np.abs(dD.cpu().numpy() - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
