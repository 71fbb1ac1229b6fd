"""Drop-in for ``train_viscosity.build_model`` (train_viscosity.py:139-231): same positional / keyword
arguments and defaults; returns an object with ``predict`` on the B200 kernels."""
from .model import MPNNModel, make_spec


def build_model(atom_vocab_size, bond_vocab_size, atom_dim=32, bond_dim=8, fp_size=32, mixing_size=20, num_steps=4,
                device="cuda", seed=0, precision="fp32", fused="auto"):
    return MPNNModel(make_spec("viscosity", atom_vocab_size, bond_vocab_size, atom_dim, bond_dim, fp_size, mixing_size,
                               num_steps), device=device, seed=seed, precision=precision, fused=fused)
