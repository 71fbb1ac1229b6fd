"""ionic-mpnn on B200: the reference's MPNN forward/backward hot path as hand-written sm_100a CUDA
behind a C ABI (include/imp_b200.h).  See DESIGN.md."""
__all__ = ["graph", "layers", "model", "synth", "viscosity", "melting_point"]
