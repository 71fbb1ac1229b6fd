"""ctypes binding of ``libimp_b200.so`` (include/imp_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised.
The product path never routes through ``oracle/`` or any CPU / eager-PyTorch implementation.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libimp_b200.so")

i32p = C.POINTER(C.c_int32)
f32p = C.c_void_p  # device float pointers are passed as raw addresses
vp = C.c_void_p


class ImpError(RuntimeError):
    pass


class PlanCapacityError(ImpError):
    """The tile plan of the fused forward ran out of tile records (MPNNModel.plan_slack): the call is repeated with the
    safe bound."""


class Ions(C.Structure):
    _fields_ = [("n_ions", C.c_int32), ("atom_ptr", vp), ("atom_ids", vp), ("edge_ptr", vp), ("edge_src", vp),
                ("edge_dst", vp), ("bond_ids", vp)]


class Graph(C.Structure):
    _fields_ = [("n_pairs", C.c_int32), ("n_atoms", C.c_int32), ("n_cat_atoms", C.c_int32), ("n_unique", C.c_int32),
                ("n_edges", C.c_int32), ("bond_vocab", C.c_int32), ("mol_ptr", vp), ("atom_id", vp), ("row_ptr", vp),
                ("col_src", vp), ("edge_bm", vp), ("bucket_ptr", vp), ("bucket_perm", vp)]


class CompactGraph(C.Structure):
    _fields_ = [("n_pairs", C.c_int32), ("n_atoms", C.c_int32), ("n_cat_atoms", C.c_int32), ("n_unique", C.c_int32),
                ("n_edges", C.c_int32), ("bond_vocab", C.c_int32), ("mol_ptr", vp), ("mol_eptr", vp), ("atom_w", vp),
                ("edge_w", vp)]


class GruWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("Wz", "bz", "Wr", "br", "Wh", "bh", "gamma", "beta")]


class ReadoutWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("W_fp", "b_fp", "W_mix", "b_mix")]


TC_FP16 = 1
TC_PRECISE_EPILOGUE = 2
TC_MP8 = 4
TC_F32_ZBUILD = 8
TC_TWO_THREADS_PER_ROW = 16
TC_THREE_CONTEXTS = 32
TC_WIDE_SPLIT_GRU = 64
TC_MSG_ONE_CHUNK_PER_CTA = 128
TC_WIDE_NO_CLUSTER = 256
TC_GEN3 = 512
TC_GEN4 = 1024
TC_GEN5 = 2048
TC_GEN7 = 4096
TC_GEN8 = 8192
PACK_DOUBLE_EDGES = 1
PACK_SHIFT_IDS = 2

# name -> (restype, argtypes); the CPU test suite checks that every symbol declared in the header is here
# and exported by the library.
SIGNATURES = {
    "imp_version": (C.c_int, []),
    "imp_last_error_string": (C.c_char_p, []),
    "imp_device_is_sm100": (C.c_int, []),
    "imp_pack_host": (C.c_int, [C.POINTER(Ions), C.POINTER(Ions), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.POINTER(Graph), C.c_int32]),
    "imp_pack_device_workspace_bytes": (C.c_int64, [C.c_int32]),
    "imp_pack_device": (C.c_int, [C.POINTER(Ions), C.POINTER(Ions), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.POINTER(Graph), vp, vp, vp, vp, vp, vp]),
    "imp_synth_ions": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp,
                                 vp, vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "imp_embed_atoms": (C.c_int, [vp, C.c_int32, vp, C.c_int32, C.c_int32, vp, vp]),
    "imp_bond_table": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(vp),
                                 C.POINTER(vp), vp]),
    "imp_message_agg": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, vp, vp]),
    "imp_edge_messages": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, vp, vp]),
    "imp_edge_messages_workspace_bytes": (C.c_int64, [C.c_int32]),
    "imp_edge_messages_grouped": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp]),
    "imp_edge_messages_grouped_tc32": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp]),
    "imp_edge_messages_grouped_tc32_planned": (C.c_int, [C.POINTER(Graph), vp, vp, C.c_int32, vp, vp, C.c_int32, vp, vp]),
    "imp_segment_sum_add": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp]),
    "imp_segment_sum": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp]),
    "imp_gated_update": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights), C.POINTER(GruWeights),
                                   C.c_float, vp, vp]),
    "imp_gated_update_wide_workspace_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_gated_update_wide": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights), C.POINTER(GruWeights),
                                        C.c_float, vp, vp, vp]),
    "imp_gru_pack_bytes": (C.c_int64, [C.c_int32]),
    "imp_gru_pack_bf16": (C.c_int, [C.POINTER(GruWeights), C.c_int32, vp, vp]),
    "imp_gru_pack_f16": (C.c_int, [C.POINTER(GruWeights), C.c_int32, vp, vp]),
    "imp_gated_update_tc": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_float, C.c_int32, vp, vp]),
    "imp_message_pack_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_message_pack": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]),
    "imp_edge_messages_tc_workspace_bytes": (C.c_int64, [C.c_int32]),
    "imp_edge_messages_tc": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp]),
    "imp_reduce_gated_update_tc": (C.c_int, [C.POINTER(Graph), vp, vp, C.c_int32, vp, vp, C.c_float, C.c_int32, vp, vp]),
    "imp_embed_atoms16": (C.c_int, [vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]),
    "imp_edge_messages_tc16": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, C.c_int32, vp, vp, vp]),
    "imp_edge_messages_tc16_plan_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_edge_messages_tc16_plan": (C.c_int, [C.POINTER(Graph), vp, vp]),
    "imp_edge_messages_tc16_planned": (C.c_int, [C.POINTER(Graph), vp, vp, C.c_int32, vp, vp, C.c_int32, vp, vp]),
    "imp_reduce_gated_update_tc16": (C.c_int, [C.POINTER(Graph), vp, vp, C.c_int32, vp, vp, C.c_float, C.c_int32, vp, vp, vp]),
    "imp_global_sum_pool": (C.c_int, [vp, vp, C.c_int32, vp, C.c_int32, vp, vp]),
    "imp_pool_head_visc": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(ReadoutWeights),
                                     C.POINTER(ReadoutWeights), vp, vp, vp, vp, vp, vp]),
    "imp_pool_head_mp": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(ReadoutWeights), C.POINTER(ReadoutWeights), vp, vp, vp, vp, vp, vp, vp]),
    "imp_fused_pack_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_fused_pack": (C.c_int, [vp, C.POINTER(GruWeights), C.c_int32, C.c_int32, C.c_int32, vp, vp]),
    "imp_mpnn_forward_fused": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_float,
                                         C.c_int32, C.c_int32, vp, vp, vp]),
    "imp_mpnn_forward_fused_compact": (C.c_int, [C.POINTER(CompactGraph), vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp,
                                                 C.c_float, C.c_int32, C.c_int32, vp, vp, vp]),
    "imp_fused_pack_planned_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_fused_pack_planned": (C.c_int, [vp, C.POINTER(GruWeights), C.c_int32, C.c_int32, vp, vp]),
    "imp_fused_pack_planned7": (C.c_int, [vp, C.POINTER(GruWeights), C.c_int32, C.c_int32, vp, vp]),
    "imp_fused_plan_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "imp_fused_plan": (C.c_int, [C.POINTER(Graph), C.POINTER(CompactGraph), C.c_int32, C.c_int32, vp, C.c_int64, vp]),
    "imp_fused_plan_compact16": (C.c_int, [C.POINTER(CompactGraph), C.c_int32, C.c_int32, vp, C.c_int64, vp]),
    "imp_mpnn_forward_fused_planned": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, vp, C.c_int32,
                                                 C.c_int32, C.c_int32, vp, C.c_float, C.c_int32, vp, vp]),
    "imp_wide_pack_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_wide_pack": (C.c_int, [vp, C.POINTER(GruWeights), C.c_int32, C.c_int32, vp, vp]),
    "imp_wide_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_mpnn_forward_wide": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_float,
                                        C.c_int32, vp, vp, vp]),
    "imp_wide_embed": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, C.c_int32, vp, vp]),
    "imp_wide_message": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp, vp]),
    "imp_wide_gates": (C.c_int, [C.POINTER(Graph), C.c_int32, vp, vp, C.c_int32, vp, vp]),
    "imp_wide_candidate": (C.c_int, [C.POINTER(Graph), C.c_int32, vp, vp, C.c_float, C.c_int32, vp, vp]),
    "imp_wide_gated_update": (C.c_int, [C.POINTER(Graph), C.c_int32, vp, vp, C.c_float, C.c_int32, vp, vp]),
    "imp_wide_pool": (C.c_int, [C.POINTER(Graph), C.c_int32, vp, vp, vp]),
    "imp_readout_visc": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(ReadoutWeights),
                                   C.POINTER(ReadoutWeights), vp, vp, vp, vp, vp, vp]),
    "imp_readout_mp": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(ReadoutWeights),
                                 C.POINTER(ReadoutWeights), vp, vp, vp, vp, vp, vp, vp]),
    "imp_bond_table_train": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(vp),
                                       C.POINTER(vp), vp]),
    "imp_readout_bwd_workspace_floats": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "imp_readout_bwd": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(ReadoutWeights),
                                  C.POINTER(ReadoutWeights), vp, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp, vp, vp, vp]),
    "imp_pool_bwd": (C.c_int, [vp, vp, C.c_int32, vp, C.c_int32, vp, vp]),
    "imp_gated_update_bwd_workspace_floats": (C.c_int64, [C.c_int32]),
    "imp_gated_update_bwd": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights),
                                       C.POINTER(GruWeights), C.c_float, vp, vp, vp, vp, vp, vp]),
    "imp_gated_update_train": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights), C.POINTER(GruWeights),
                                         C.c_float, vp, vp, vp, vp, vp]),
    "imp_gated_update_tc32": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights), C.POINTER(GruWeights),
                                        C.c_float, vp, vp, vp, vp, vp]),
    "imp_gated_update_bwd_stored": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights),
                                              C.POINTER(GruWeights), C.c_float, vp, vp, vp, vp, vp, vp]),
    "imp_gated_update_bwd_tc": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(GruWeights),
                                          C.POINTER(GruWeights), C.c_float, vp, vp, vp, vp, vp, vp]),
    "imp_message_agg_bwd": (C.c_int, [C.POINTER(Graph), vp, C.c_int32, vp, vp, vp, vp]),
    "imp_bond_transform_bwd": (C.c_int, [C.POINTER(Graph), vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp,
                                         vp, vp, vp, vp, vp, vp]),
    "imp_embed_bwd_workspace_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_embed_bwd": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]),
    "imp_clip_adam": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_int32, vp]),
    "imp_clip_adam_sparse": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_float, vp, vp]),
    "imp_eval_loss": (C.c_int, [vp, vp, C.c_int64, vp, vp, vp, C.c_int32, vp, vp, vp]),
    "imp_sumsq": (C.c_int, [vp, C.c_int64, vp, vp, vp]),
    "imp_occ_pack_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "imp_occ_pack": (C.c_int, [vp, C.c_int32, C.c_int32, vp, vp]),
    "imp_bond_occurrence_norm2_workspace_floats": (C.c_int64, [C.c_int32]),
    "imp_bond_occurrence_norm2": (C.c_int, [C.POINTER(Graph), C.c_int32, vp, C.c_int32, C.POINTER(vp), C.POINTER(vp), C.c_int32,
                                            C.c_int32, vp, vp, vp, vp, vp]),
    "imp_dense": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp, vp]),
    "imp_dense_bwd": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, vp, vp]),
    "imp_batchnorm": (C.c_int, [vp, C.c_int64, C.c_int32, vp, vp, vp, vp, C.c_float, C.c_float, C.c_int32, vp, vp, vp, vp]),
    "imp_batchnorm_bwd": (C.c_int, [vp, vp, C.c_int64, C.c_int32, vp, vp, vp, vp, vp, vp, vp]),
    "imp_dropout": (C.c_int, [vp, C.c_int64, C.c_float, C.c_uint64, vp, vp]),
    "imp_huber": (C.c_int, [vp, vp, C.c_int64, C.c_float, C.c_float, vp, vp, vp]),
    "imp_add": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
    "imp_tc_selftest": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
}

_lib = None


def load():
    """Loads the library (once).  Raises ImpError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImpError(f"{LIB_PATH} is missing: run `python -m ionic_mpnn_b200.build` (or __graft_entry__.build()). "
                       "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.imp_version() < 100:
        raise ImpError("libimp_b200.so is older than the Python package")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().imp_last_error_string().decode(errors="replace")
        kind = "CUDA error" if rc > 0 else "argument error"
        raise ImpError(f"{what}: {kind} {rc}: {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
