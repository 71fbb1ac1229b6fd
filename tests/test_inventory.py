"""CPU: the weight inventory of the host-side models equals the reference's (variables / parameters of the
graphs built by train_viscosity.py:139-231 and train_melting_point.py:137-215), and the oracle agrees."""
import numpy as np
import pytest

from ionic_mpnn_b200.model import make_spec, param_shapes


@pytest.mark.parametrize("kind,kw,n_vars,n_params", [
    ("viscosity", {}, 84, 124_007),
    ("melting_point", {}, 86, 8_520_873),
    ("viscosity", {"atom_dim": 256, "num_steps": 6}, 120, 11_075_559),
])
def test_weight_inventory_matches_the_reference(kind, kw, n_vars, n_params):
    from oracle import ref_model

    shapes = param_shapes(make_spec(kind, **kw))
    assert len(shapes) == n_vars
    assert sum(int(np.prod(s)) for s in shapes.values()) == n_params
    ospec = ref_model.make_spec(kind, **kw)
    oshapes = ref_model.param_shapes(ospec)
    assert {k: tuple(v) for k, v in oshapes.items()} == {k: tuple(v) for k, v in shapes.items()}
    # layouts the kernels rely on: bond_transform (K, d, d); Dense kernels (2d, d); embeddings (vocab, dim)
    s = make_spec(kind, **kw)
    d, K = s["atom_dim"], s["bond_dim"]
    assert shapes["cat_bmm_0.bond_transform"] == (K, d, d)
    assert shapes["an_gu_0.dense_h.kernel"] == (2 * d, d)
    assert shapes["atom_emb"] == (124, d) and shapes["bond_emb"] == (72, K)
