"""Small invocations of the round-2 kernel paths, meant to be run under compute-sanitizer --tool memcheck (debug aid):
tile plan (all three feeds, degenerate molecules), planned forward generations 5 and 6, readout, tensor-core backward,
transfer head."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionic_mpnn_b200 import _lib, graph, synth  # noqa: E402
from ionic_mpnn_b200.model import MPNNModel, make_spec  # noqa: E402
from ionic_mpnn_b200.transfer import TransferModel, UNFREEZE_KEYS  # noqa: E402

for n, lo, hi in ((1, 10, 40), (130, 1, 128), (700, 10, 40)):
    b, _, _ = graph.synth_batch(n, seed=3 + n, n_min=lo, n_max=hi)
    for flags in (0, _lib.TC_GEN5):
        m = MPNNModel(make_spec("viscosity"), seed=1, precision="fp16", fused=True)
        m.extra_tc_flags = flags
        y = m.predict(b)
        yc = m.forward_packed(b.to_compact("cuda")).cpu().numpy()
        assert np.isfinite(y).all() and np.array_equal(y.reshape(-1), yc)
        if b.narrow_ok:
            yn = m.forward_packed(b.to_compact("cuda", narrow=True)).cpu().numpy()
            assert np.array_equal(yc, yn)
    print("planned forward", n, lo, hi, "ok", flush=True)
out, _ = MPNNModel(make_spec("viscosity"), seed=1, precision="fp16", fused=True).predict_stream([graph.synth_batch(300, seed=s)[0] for s in (1, 2, 3)])
torch.cuda.synchronize()
assert torch.isfinite(out).all()
print("predict_stream ok", flush=True)
recs = synth.make_records(70, seed=3, label="log_eta")
tb = graph.pack_records(recs, 72, label="log_eta")
for tcb in (True, False):
    m = MPNNModel(make_spec("viscosity"), seed=2)
    m.tc_backward = tcb
    l0 = float(m.train_step(tb))
    l1 = float(m.train_step(tb))
    assert l1 < l0
    print("train ok", tcb, l0, l1, flush=True)
recs = synth.make_records(40, seed=4, label="mp")
tb = graph.pack_records(recs, 72, label="mp")
tb.target = np.random.default_rng(0).normal(size=40).astype(np.float32)
tb.to("cuda")
tm = TransferModel(MPNNModel(make_spec("viscosity"), seed=2))
a = float(tm.compile().train_step(tb))
tm.unfreeze(UNFREEZE_KEYS).compile()
b2 = float(tm.train_step(tb, lr=1e-4))
assert np.isfinite([a, b2]).all() and np.isfinite(tm.predict(recs)).all()
print("transfer ok", a, b2, flush=True)
