"""GPU (B200): the 3xTF32 bucket-grouped message kernel (imp_edge_messages_grouped_tc32, csrc/msg_tc32.cu) against the exact
fp32 SIMT kernel (imp_edge_messages_grouped), forward and transposed."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(batch, vb, seed):
    from ionic_mpnn_b200 import _lib

    rng = np.random.default_rng(seed)
    d = 32
    batch.to("cuda")
    g = batch.c_struct()
    x = torch.from_numpy(rng.normal(size=(batch.n_atoms, d)).astype(np.float32)).cuda()
    tabs = [torch.from_numpy((rng.normal(size=(vb, d, d)) * 0.3).astype(np.float32)).cuda() for _ in range(2)]
    ws = torch.zeros(_lib.load().imp_edge_messages_workspace_bytes(vb) // 4 + 8, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    lib = _lib.load()
    plan = torch.empty(max(int(lib.imp_edge_messages_tc16_plan_bytes(batch.n_unique, vb)), 16), dtype=torch.uint8, device="cuda")
    _lib.call("imp_edge_messages_tc16_plan", C.byref(g), plan.data_ptr(), st)
    for tr in (0, 1, 2, 3):  # bit 0: transposed; bit 1: operand terms rounded to tf32 (else truncated: the training form)
        msg = torch.full((batch.n_unique, d), 7.0, device="cuda")
        _lib.call("imp_edge_messages_grouped_tc32_planned", C.byref(g), plan.data_ptr(), x.data_ptr(), d, tabs[0].data_ptr(),
                  tabs[1].data_ptr(), tr, msg.data_ptr(), st)
        torch.cuda.synchronize()
        out[("planned", tr)] = msg.cpu().numpy().astype(np.float64)
    for name in ("imp_edge_messages_grouped", "imp_edge_messages_grouped_tc32"):
        for tr in (0, 1):
            msg = torch.full((batch.n_unique, d), 7.0, device="cuda")
            _lib.call(name, C.byref(g), x.data_ptr(), d, tabs[0].data_ptr(), tabs[1].data_ptr(), tr, msg.data_ptr(), ws.data_ptr(), st)
            torch.cuda.synchronize()
            out[(name, tr)] = msg.cpu().numpy().astype(np.float64)
    return out


@pytest.mark.parametrize("n_pairs,seed,lo,hi,vb", [(1, 1, 10, 40, 72), (300, 2, 10, 40, 72), (5000, 3, 10, 40, 72), (200, 4, 1, 128, 7)])
def test_tc32_messages_match_the_fp32_kernel(n_pairs, seed, lo, hi, vb):
    from ionic_mpnn_b200 import graph

    if vb == 72:
        batch, _, _ = graph.synth_batch(n_pairs, seed=seed, n_min=lo, n_max=hi)
    else:  # a small bond vocabulary with unused types (empty buckets), tiny cations / large anions
        cat = graph.synth_flat(n_pairs, 71, 2, 4, atom_types=9, bond_types=vb - 2)
        an = graph.synth_flat(n_pairs, 72, lo, hi, atom_types=9, bond_types=vb - 2)
        batch = graph.pack_flat(cat, an, vb)
    out = _run(batch, vb, seed)
    for tr in (0, 1):
        a, b = out[("imp_edge_messages_grouped", tr)], out[("imp_edge_messages_grouped_tc32", tr)]
        assert np.isfinite(b).all()
        err = np.abs(a - b).max() / max(np.abs(a).max(), 1.0)
        print(tr, f"{err:.2e}")
        assert err <= 2e-6, (tr, err)  # fp32-class: summation order and the 3xTF32 split (~2^-21 per product)
        assert np.array_equal(out[("planned", tr | 2)], b), "the planned persistent kernel (rounded terms) computes the same rows"
        err_t = np.abs(a - out[("planned", tr)]).max() / max(np.abs(a).max(), 1.0)
        assert err_t <= 4e-6, (tr, err_t)  # truncated terms (training form): twice as coarse


def test_tc32_messages_are_bit_reproducible():
    from ionic_mpnn_b200 import graph

    batch, _, _ = graph.synth_batch(2000, seed=11)
    a, b = _run(batch, 72, 5), _run(batch, 72, 5)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
