import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from ionic_mpnn_b200 import _lib
rng = np.random.default_rng(0)
def tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
for N, K in ((64, 8), (96, 64)):
    At = tf32(rng.standard_normal((K, 128)).astype(np.float32)); Bt = tf32(rng.standard_normal((K, N)).astype(np.float32))
    for swap in (6, 2, 3, 4, 5, 0, 1):
        dA, dB = torch.from_numpy(At).cuda(), torch.from_numpy(Bt).cuda()
        dD = torch.zeros(128, N, dtype=torch.float32, device="cuda")
        _lib.call("imp_tc_selftest", dA.data_ptr(), dB.data_ptr(), dD.data_ptr(), N, K, 4, swap, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        want = At.astype(np.float64).T @ Bt.astype(np.float64)
        err = np.abs(dD.cpu().numpy() - want).max() / max(1.0, np.abs(want).max())
        print(f"N={N} K={K} swap={swap}: rel err {err:.3e}", flush=True)
