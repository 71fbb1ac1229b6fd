"""Drop-in for ``train_melting_point.build_model`` (train_melting_point.py:137-215): no ``bond_dim``
argument, the bond embedding is ``atom_dim**2`` wide (:146)."""
from .model import MPNNModel, make_spec


def build_model(atom_vocab_size, bond_vocab_size, atom_dim=32, fp_size=32, mixing_size=20, num_steps=4, device="cuda",
                seed=0, precision="fp32", fused="auto"):
    return MPNNModel(make_spec("melting_point", atom_vocab_size, bond_vocab_size, atom_dim, None, fp_size, mixing_size,
                               num_steps), device=device, seed=seed, precision=precision, fused=fused)
