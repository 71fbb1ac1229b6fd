/*
 * imp_b200.h -- C ABI of the B200-native (sm_100a) ionic-mpnn MPNN hot path.
 *
 * The reference (goalheart/ionic-mpnn) has no native interface: the path sits behind the Keras
 * Layer protocol of models/layers.py.  Each entry point below names the reference code it
 * replaces (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns int: 0 = OK, >0 = a cudaError_t, <0 = IMP_ERR_* (argument / shape);
 *     imp_last_error_string() describes the last failure on the calling thread.
 *   - no allocation and no exceptions cross the boundary: the caller owns every buffer and passes raw
 *     pointers + sizes.  "d_" = device pointer, "h_" = host pointer.
 *   - kernels are enqueue-only on `stream` (a cudaStream_t passed as void*); no hidden synchronisation;
 *     re-entrant across streams.  The device is the caller's current device.
 *   - float tensors are row-major fp32 unless a name says bf16; index tensors are int32.
 *   - atoms of a batch are tower-major (all cation atoms, then all anion atoms); `n_cat_atoms` splits
 *     them.  Per-tower weights are passed as two pointers / two structs, [0] = cation, [1] = anion
 *     (the reference instantiates separate layers per tower and step: train_viscosity.py:176-189).
 */
#ifndef IMP_B200_H
#define IMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMP_VERSION 100

#define IMP_ERR_ARG (-1)         /* null pointer / negative size / inconsistent sizes   */
#define IMP_ERR_DIM (-2)         /* atom_dim / fp / mix size not supported by the kernels */
#define IMP_ERR_INDEX (-3)       /* edge endpoint or id out of range (host packer)       */
#define IMP_ERR_CAPACITY (-4)    /* caller-provided output capacity too small            */
#define IMP_ERR_UNSUPPORTED (-5) /* needs sm_100a / feature not built                     */

#define IMP_TILE_M 128 /* row tile of the bond-bucketed message GEMM */

int imp_version(void);
const char* imp_last_error_string(void);
/* 1 if the current device is sm_100 (B200); the tcgen05 entry points refuse to run otherwise. */
int imp_device_is_sm100(void);

/* ---------------------------------------------------------------------------------------------
 * Graph batch.  Replaces the padded tensors of build_inputs (train_viscosity.py:291-314):
 * pad_sequences_1d (:52-59) + preprocess_edges_and_bonds (:76-110) + the masks applied later in
 * BondMatrixMessage (models/layers.py:114-115) and Reduce (models/layers.py:74-76).
 * Bit-exact specification: oracle/ref_pack.py.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_ions;           /* = number of pairs                                              */
  const int32_t* atom_ptr;  /* [n_ions+1] offsets into atom_ids                               */
  const int32_t* atom_ids;  /* vocabulary ids                                                  */
  const int32_t* edge_ptr;  /* [n_ions+1] offsets into edge_src / edge_dst / bond_ids          */
  const int32_t* edge_src;  /* ion-local, 0-based (src/featurize.py:60-63 emits both directions) */
  const int32_t* edge_dst;
  const int32_t* bond_ids;
} imp_ions_t;

typedef struct {
  int32_t n_pairs;
  int32_t n_atoms;     /* N: atoms of both towers                               */
  int32_t n_cat_atoms; /* atoms [0, n_cat_atoms) belong to the cation tower     */
  int32_t n_unique;    /* Eu: unique live (dst, bond, src) entries              */
  int32_t n_edges;     /* E: live entries counted with multiplicity */
  int32_t bond_vocab;  /* V_b (embedding rows, incl. the padding row 0)         */
  int32_t* mol_ptr;     /* [2*n_pairs+1]                                         */
  int32_t* atom_id;     /* [N] shifted ids; id 0 = "not pooled" (models/layers.py:163) */
  int32_t* row_ptr;     /* [N+1] CSR over destination atoms                      */
  int32_t* col_src;     /* [Eu] global source atom; rows sorted by (bond, src)   */
  int32_t* edge_bm;     /* [Eu] bond id | multiplicity << 16                     */
  int32_t* bucket_ptr;  /* [2*V_b+1] entries grouped by (tower, bond)            */
  int32_t* bucket_perm; /* [Eu] entry indices in bucket order (stable)           */
} imp_graph_t;

/* flags for imp_pack_host */
#define IMP_PACK_DOUBLE_EDGES 1 /* append the reverse of every entry (train_viscosity.py:87-91)      */
#define IMP_PACK_SHIFT_IDS 2    /* +1 on atom and bond ids (train_viscosity.py:255-262)             */

/* Host packer.  `out` arrays are caller-allocated: mol_ptr[2P+1], atom_id[N], row_ptr[N+1],
 * bucket_ptr[2*V_b+1], and col_src / edge_bm / bucket_perm with capacity `edge_capacity` entries
 * (an upper bound is the number of input entries, doubled when IMP_PACK_DOUBLE_EDGES).  max_edges < 0
 * disables the reference's truncation to 2*max_edges entries per ion.  Counts are returned in `out`. */
int imp_pack_host(const imp_ions_t* cation, const imp_ions_t* anion, int32_t bond_vocab, int32_t max_edges,
                  int32_t flags, int32_t edge_capacity, imp_graph_t* out, int32_t n_threads);

/* Device packer (same contract and bit-identical mol_ptr / atom_id / row_ptr / col_src / edge_bm as imp_pack_host; the
 * bond-bucket permutation of the staged message / training kernels is NOT produced).  `cation` / `anion` hold DEVICE
 * pointers; n_cat_atoms / n_atoms are the host-known totals (atom_ptr[P] of each tower).  Optional compact-feed outputs
 * (d_mol_eptr[2P+1], d_atom_w[N], d_edge_w[capacity]; see imp_compact_graph_t) may be NULL.  Enqueue-only: the four
 * int32 of d_counts receive n_unique, a status (0 or an IMP_ERR_* code: index out of range, a molecule with more than
 * 512 doubled entries or 1024 atoms, capacity exceeded) and the 64-bit sum of multiplicities; the caller reads them
 * back after synchronising.  d_workspace: imp_pack_device_workspace_bytes(n_pairs). */
int64_t imp_pack_device_workspace_bytes(int32_t n_pairs);
int imp_pack_device(const imp_ions_t* cation, const imp_ions_t* anion, int32_t n_cat_atoms, int32_t n_atoms,
                    int32_t bond_vocab, int32_t max_edges, int32_t flags, int32_t edge_capacity, imp_graph_t* out,
                    int32_t* d_mol_eptr, uint16_t* d_atom_w, uint32_t* d_edge_w, int32_t* d_counts, void* d_workspace,
                    void* stream);

/* Benchmark-sized synthetic ions (recipe of SURVEY 8d; splitmix64 stream, see synth.py).  Two-phase:
 * call with all output pointers NULL to get *n_atoms / *n_entries, then again with buffers. */
int imp_synth_ions(uint64_t seed, int32_t n_ions, int32_t n_min, int32_t n_max, int32_t atom_types,
                   int32_t bond_types, int32_t skewed, int32_t* atom_ptr, int32_t* atom_ids, int32_t* edge_ptr,
                   int32_t* edge_src, int32_t* edge_dst, int32_t* bond_ids, int64_t* n_atoms, int64_t* n_entries);

/* ---------------------------------------------------------------------------------------------
 * K1  Embedding(atom)                        train_viscosity.py:163,171 ; train_melting_point.py:149,157
 *     d_h0[v,:] = d_atom_emb[atom_id[v],:]
 * ------------------------------------------------------------------------------------------- */
int imp_embed_atoms(const float* d_atom_emb, int32_t atom_vocab, const int32_t* d_atom_id, int32_t n_atoms,
                    int32_t d, float* d_h0, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  bond-matrix table.  Replaces Embedding(bond) + the per-edge tf.tensordot of
 *     BondMatrixMessage.call (models/layers.py:108): table[v,l,m] = sum_k bond_emb[v,k] * W[k,l,m],
 *     once per (tower, step) instead of once per padded edge.  `n_tables` weight tensors are processed
 *     in one launch.  Outputs (either may be NULL):
 *       d_table[i]     [V_b, d, d]  row-major, A[l,m] as in the reference
 *       d_table_il[i]  [V_b, d/4, d, 4]  "lane-interleaved" copy read by imp_message_agg
 * ------------------------------------------------------------------------------------------- */
#define IMP_MAX_TABLES 32
int imp_bond_table(const float* d_bond_emb, int32_t bond_vocab, int32_t bond_dim, int32_t d, int32_t n_tables,
                   const float* const* h_W /* host array of n_tables device pointers [K,d,d] */,
                   float* const* h_table, float* const* h_table_il, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3+K4 fused  BondMatrixMessage o Reduce (models/layers.py:100-117 + :57-83; the fused signature of
 *     models/bond_matrix_message.py:37-65):  agg[v] = sum_{e in row v} mult_e * table[bond_e] @ h[src_e]
 *     Deterministic (fixed CSR order), no message tensor is materialised.
 * ------------------------------------------------------------------------------------------- */
int imp_message_agg(const imp_graph_t* g /* host struct, device arrays */, const float* d_h, int32_t d,
                    const float* d_table_il_cat, const float* d_table_il_an, float* d_agg, void* stream);

/* K3  BondMatrixMessage.call (models/layers.py:100-117): per-entry messages in CSR order,
 *     msg[e] = mult_e * table[bond_e] @ h[src_e], computed bucket by bucket (tower, bond). */
int imp_edge_messages(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_cat,
                      const float* d_table_an, float* d_msg /* [Eu, d] */, void* stream);

/* K3 bucket-grouped for atom_dim 32 (exact fp32): one (tower, bond) bucket chunk per CTA, the 4 KB bond matrix staged in
 * shared memory and shared by the chunk's entries.  transposed = 0: msg[e] = mult * T[b] x[src_e] (the forward layer);
 * transposed = 1: msg[e] = mult * T[b]^T x[src_e] (its backward with respect to the atom states: the live entry set is
 * symmetric, so dh = segment sum of these rows).  d_workspace: imp_edge_messages_workspace_bytes(bond_vocab) bytes. */
int64_t imp_edge_messages_workspace_bytes(int32_t bond_vocab);
int imp_edge_messages_grouped(const imp_graph_t* g, const float* d_x, int32_t d, const float* d_table_cat,
                              const float* d_table_an, int32_t transposed, float* d_msg /* [Eu, d] */, void* d_workspace,
                              void* stream);
/* The same contract on the tensor cores (csrc/msg_tc32.cu): source rows and T[b] split into two tf32 terms, three tcgen05.mma
 * kind::tf32 per K step (fp32-class accuracy, ~2^-21 relative per product); a gather / scatter stream instead of 1,024 FFMA per
 * entry.  atom_dim 32. */
int imp_edge_messages_grouped_tc32(const imp_graph_t* g, const float* d_x, int32_t d, const float* d_table_cat,
                              const float* d_table_an, int32_t transposed, float* d_msg /* [Eu, d] */, void* d_workspace,
                              void* stream);
/* Planned, persistent form (the training step's): d_plan = the per-batch index plan of imp_edge_messages_tc16_plan
 * (imp_edge_messages_tc16_plan_bytes bytes); source rows arrive by cp.async one chunk ahead of the MMAs.
 * transposed: bit 0 = T[b]^T (backward); bit 1 = operand terms rounded to tf32 instead of truncated (the fp32 inference route:
 * 4 % slower, the lo terms unbiased). */
int imp_edge_messages_grouped_tc32_planned(const imp_graph_t* g, const void* d_plan, const float* d_x, int32_t d,
                                           const float* d_table_cat, const float* d_table_an, int32_t transposed,
                                           float* d_msg /* [Eu, d] */, void* stream);

/* K4  Reduce.call (models/layers.py:57-83): agg[v] = sum of msg rows row_ptr[v]..row_ptr[v+1] (imp_segment_sum_add: added
 * to the rows already in d_agg -- the message term of dL/dh in the backward pass). */
int imp_segment_sum(const imp_graph_t* g, const float* d_msg, int32_t d, float* d_agg, void* stream);
int imp_segment_sum_add(const imp_graph_t* g, const float* d_msg, int32_t d, float* d_agg, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5  GatedUpdate.call (models/layers.py:142-156): z, r gates, candidate, blend, LayerNormalization
 *     (eps = 1e-3, biased variance), residual.  Dense kernels are (2d, d): rows [0,d) multiply h (or
 *     r*h), rows [d,2d) multiply agg (concat order, models/layers.py:144,150).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float *Wz, *bz, *Wr, *br, *Wh, *bh, *gamma, *beta;
} imp_gru_weights_t;

int imp_gated_update(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                     const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out,
                     void* stream);

/* K5 for wide atom states (atom_dim a multiple of 16, e.g. 128 / 256 -- the "wide/deep" variant): three tiled fp32
 * GEMMs with fused gate epilogues + one LayerNorm/residual kernel.  d_workspace: imp_gated_update_wide_workspace_floats. */
int64_t imp_gated_update_wide_workspace_floats(int32_t n_atoms, int32_t d);
int imp_gated_update_wide(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                          const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out,
                          float* d_workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6  GlobalSumPool (models/layers.py:161-164) + Dense(fp, relu) (train_viscosity.py:189) +
 *     Dense(mix, relu) x2 + AddTwoTensors (:197-201) + head:
 *       viscosity   Dense(3) -> A, B = clip(softplus,0,20), C = clip(softplus,0.1,50),
 *                   log_eta = A + B / (T/100 + C + 1e-6)      (:204-214, models/layers.py:10-42)
 *       melting pt  Dense(fp2, relu) -> Dense(1)               (train_melting_point.py:191-198)
 * ------------------------------------------------------------------------------------------- */
/* K5 on the tensor cores (tcgen05, 16-bit operands, fp32 accumulation in TMEM; the "2e-2" path).  Weights are
 * packed once per weight update into the UMMA operand layout: imp_gru_pack_bytes(d) bytes per (tower, step).
 * flags (IMP_TC_*, the same for pack and update): IMP_TC_FP16 = IEEE-half operands instead of bfloat16;
 * IMP_TC_PRECISE_EPILOGUE = expf / tanhf / sqrtf in the gate epilogue instead of tanh.approx / rsqrt. */
#define IMP_TC_FP16 1
#define IMP_TC_PRECISE_EPILOGUE 2
#define IMP_TC_MP8 4
#define IMP_TC_F32_ZBUILD 8
#define IMP_TC_TWO_THREADS_PER_ROW 16
#define IMP_TC_THREE_CONTEXTS 32
#define IMP_TC_WIDE_SPLIT_GRU 64
#define IMP_TC_WIDE_NO_CLUSTER 256 /* wide GatedUpdate without the 2-CTA weight multicast (comparison) */
#define IMP_TC_MSG_ONE_CHUNK_PER_CTA 128 /* imp_edge_messages_tc16: the non-pipelined kernel (comparison) */
#define IMP_TC_GEN3 512 /* imp_mpnn_forward_fused: kept for callers of round 1; the self-contained kernel IS generation 3 */
#define IMP_TC_GEN4 1024 /* imp_mpnn_forward_fused: the fourth-generation kernel (arrive-and-continue; measured slower) */
#define IMP_TC_GEN5 2048 /* imp_mpnn_forward_fused_planned: the fifth-generation kernel (weights from imp_fused_pack; comparison) */
#define IMP_TC_GEN7 4096 /* imp_mpnn_forward_fused_planned: the seventh-generation kernel (weights from imp_fused_pack_planned7) */
#define IMP_TC_GEN8 8192 /* imp_mpnn_forward_fused_planned: the eighth-generation kernel (weights from imp_fused_pack_planned7) */
int64_t imp_gru_pack_bytes(int32_t d);
int imp_gru_pack_bf16(const imp_gru_weights_t* w, int32_t d, void* d_packed, void* stream);
int imp_gru_pack_f16(const imp_gru_weights_t* w, int32_t d, void* d_packed, void* stream);
int imp_gated_update_tc(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                        const void* d_packed_cat, const void* d_packed_an, float eps, int32_t flags,
                        float* d_h_out, void* stream);

/* K3 on the tensor cores for atom_dim 32, any bond_dim (the message kernel of the melting-point model, bond_dim 1024):
 * the bond-type-grouped GEMM  M_b = H_src,b . T[b]^T  per (tower, bond) bucket -- chunks of 128 gathered source rows as
 * the shared-memory A operand, T[b] as B, tcgen05.mma with fp32 accumulators in TMEM, rows scaled by the multiplicity
 * and written at the entry's CSR position (csrc/msg_tc.cu).  Same output contract as imp_edge_messages at 16-bit
 * operand precision (the "2e-2" path); follow with imp_segment_sum (Reduce).  Needs the graph's bucket arrays.
 *   imp_message_pack   table [V_b, d, d] (imp_bond_table's d_table) -> imp_message_pack_bytes() bytes, once per weight
 *                      update and per (tower, step);  flags: IMP_TC_FP16 (must match the forward call).
 *   d_workspace        imp_edge_messages_tc_workspace_bytes(bond_vocab) bytes (per-bucket chunk offsets). */
int64_t imp_message_pack_bytes(int32_t bond_vocab, int32_t d);
int imp_message_pack(const float* d_table, int32_t bond_vocab, int32_t d, int32_t flags, void* d_packed, void* stream);
int64_t imp_edge_messages_tc_workspace_bytes(int32_t bond_vocab);
int imp_edge_messages_tc(const imp_graph_t* g, const float* d_h, int32_t d, const void* d_packed_cat, const void* d_packed_an,
                         int32_t flags, float* d_msg /* [Eu, d] */, void* d_workspace, void* stream);

/* Reduce.call (models/layers.py:57-83) folded into the load stage of the tensor-core GatedUpdate: d_msg holds the
 * per-entry message rows [Eu, d] in CSR order (imp_edge_messages / imp_edge_messages_tc); each atom's rows are summed
 * in entry order while they are staged, so agg is never written.  Bit-identical to imp_segment_sum + imp_gated_update_tc. */
int imp_reduce_gated_update_tc(const imp_graph_t* g, const float* d_h, const float* d_msg, int32_t d,
                               const void* d_packed_cat, const void* d_packed_an, float eps, int32_t flags,
                               float* d_h_out, void* stream);

/* 16-bit I/O forms of the two calls above (what the staged tensor forward runs when no intermediates are kept): the atom
 * states have an operand-format copy h16 [N, d] (IEEE half or bfloat16 per IMP_TC_FP16) that the message kernel gathers
 * and the GatedUpdate writes next to the fp32 state; message rows are stored in the operand format too.  Per step
 * 2.3 instead of 3.2 GB at BASELINE configs[1].  imp_embed_atoms16 = Embedding(atom) writing both copies. */
int imp_embed_atoms16(const float* d_atom_emb, int32_t atom_vocab, const int32_t* d_atom_id, int32_t n_atoms, int32_t d,
                      int32_t flags, float* d_h0, void* d_h0_16, void* stream);
int imp_edge_messages_tc16(const imp_graph_t* g, const void* d_h16, int32_t d, const void* d_packed_cat,
                           const void* d_packed_an, int32_t flags, void* d_msg16 /* [Eu, d] 16-bit */, void* d_workspace,
                           void* stream);
/* Planned form of imp_edge_messages_tc16 (what the model's forward runs): the index work of a batch -- chunk offsets and
 * bucket-ordered copies of the source atoms and bond | multiplicity words -- is done once per batch into d_plan
 * (imp_edge_messages_tc16_plan_bytes bytes), and every step's kernel reads its indices with independent coalesced loads
 * two chunks ahead of their use.  bond_vocab <= 256.  Bit-identical message rows. */
int64_t imp_edge_messages_tc16_plan_bytes(int32_t n_unique, int32_t bond_vocab);
int imp_edge_messages_tc16_plan(const imp_graph_t* g, void* d_plan, void* stream);
int imp_edge_messages_tc16_planned(const imp_graph_t* g, const void* d_plan, const void* d_h16, int32_t d,
                                   const void* d_packed_cat, const void* d_packed_an, int32_t flags, void* d_msg16, void* stream);
int imp_reduce_gated_update_tc16(const imp_graph_t* g, const float* d_h, const void* d_msg16, int32_t d,
                                 const void* d_packed_cat, const void* d_packed_an, float eps, int32_t flags,
                                 float* d_h_out, void* d_h16_out, void* stream);

/* GlobalSumPool.call alone (models/layers.py:161-164): out[m,:] = sum of h rows of molecule m whose
 * atom_id > 0.  `n_mols` molecules delimited by d_mol_ptr[n_mols+1]. */
int imp_global_sum_pool(const int32_t* d_mol_ptr, const int32_t* d_atom_id, int32_t n_mols, const float* d_h, int32_t d,
                        float* d_out /* [n_mols, d] */, void* stream);

typedef struct {
  const float *W_fp, *b_fp;   /* [d, fp], [fp]   */
  const float *W_mix, *b_mix; /* [fp, mix], [mix] */
} imp_readout_weights_t;

int imp_pool_head_visc(const imp_graph_t* g, const float* d_h, int32_t d, int32_t fp, int32_t mix,
                       const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an,
                       const float* d_W_head /* [mix,3] */, const float* d_b_head /* [3] */,
                       const float* d_T /* [P] kelvin */, float* d_out /* [P] */,
                       float* d_aux /* optional [P, 2*d + 2*fp + mix + 3]: pools, fps, mixed, (A,B,C) */,
                       void* stream);

int imp_pool_head_mp(const imp_graph_t* g, const float* d_h, int32_t d, int32_t fp, int32_t mix, int32_t fp2,
                     const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an,
                     const float* d_W1 /* [mix,fp2] */, const float* d_b1, const float* d_W2 /* [fp2,1] */,
                     const float* d_b2, float* d_out /* [P] */, float* d_aux /* optional, as above w/o (A,B,C) */,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused forward (the throughput path): Embedding(atom) -> [BondMatrixMessage o Reduce -> GatedUpdate] x steps ->
 * GlobalSumPool for BOTH towers in one persistent tcgen05 kernel; atom states never leave the SM between steps.
 * Replaces train_viscosity.py:163,171-187 (the `encode` loop) + models/layers.py:57-164 for atom_dim 32,
 * bond_dim 8, steps <= 4, molecules <= 128 atoms.  Other shapes: the staged kernels above (IMP_ERR_DIM here).
 *
 *   imp_fused_pack     once per weight update and per (tower, step): bond_transform (K,d,d) and the eight
 *                      GatedUpdate variables -> one UMMA-ready block of imp_fused_pack_bytes() bytes.
 *                      d_packed of imp_mpnn_forward_fused is [2 towers][steps] such blocks, cation first.
 *   flags              IMP_TC_FP16: 16-bit operands are IEEE half (11-bit significand) instead of bfloat16;
 *                      both accumulate in fp32.  Must match between pack and forward.
 *                      IMP_TC_PRECISE_EPILOGUE: expf/tanhf/sqrtf instead of tanh.approx/rsqrt.
 *                      IMP_TC_F32_ZBUILD: with IMP_TC_FP16, accumulate the Z rows in fp32 (first-generation kernel)
 *                      instead of packed half2 (default; 2 products per lane-instruction, rows sorted by degree).
 *                      IMP_TC_MP8: smaller register tile in the fp32 Z build (tuning switch, same results).
 *                      IMP_TC_TWO_THREADS_PER_ROW: second-generation half kernel (2 contexts x 256 threads per SM) instead
 *                      of the default third generation (4 contexts x 128 threads, blocking context barriers;
 *                      IMP_TC_THREE_CONTEXTS: its 3-context form); IMP_TC_GEN4: fourth generation (csrc/fused_fwd4.cu:
 *                      arrive-and-continue synchronisation, r*h fold, 3-instruction LayerNorm; 8 % slower, kept for comparison).
 *                      Pack and forward must be called with the same flags (the Wc block layout differs).
 *                      The fastest form is the PLANNED forward below (fifth generation), which reads a tile plan.
 *   max_mol_atoms      largest molecule of the batch (the caller knows it from mol_ptr); > 128 is refused.
 *   d_pooled           [2 * n_pairs, d] molecule sums in mol_ptr order -> imp_readout_visc / imp_readout_mp.
 *   d_status           optional device int, set to 1 if the kernel met a molecule that does not fit a tile.
 * ------------------------------------------------------------------------------------------- */
int64_t imp_fused_pack_bytes(int32_t d, int32_t bond_dim);
int imp_fused_pack(const float* d_bond_transform /* [K,d,d] */, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                   int32_t flags, void* d_packed, void* stream);
int imp_mpnn_forward_fused(const imp_graph_t* g, const float* d_atom_emb, int32_t atom_vocab, const float* d_bond_emb,
                           int32_t d, int32_t bond_dim, int32_t steps, const void* d_packed, float eps, int32_t flags,
                           int32_t max_mol_atoms, float* d_pooled, int32_t* d_status, void* stream);

/* Compact input feed of the fused forward (halves the host->device bytes of a streamed sweep: 0.49 instead of 1.15 KB
 * per pair).  Same graph as imp_graph_t, for vocabularies <= 256, in-degrees <= 255, molecules <= 256 atoms:
 *   atom_w[v]  = atom_id | in_degree << 8                         (replaces atom_id[] and row_ptr[])
 *   edge_w[e]  = src (index inside its molecule) | bond << 8 | multiplicity << 16   (replaces col_src[] and edge_bm[])
 *   mol_eptr[m] = first CSR entry of molecule m (= row_ptr[mol_ptr[m]])
 * Built on the host by PackedGraphBatch.compact() (graph.py); results are bit-identical to imp_mpnn_forward_fused. */
typedef struct {
  int32_t n_pairs, n_atoms, n_cat_atoms, n_unique, n_edges, bond_vocab;
  const int32_t* mol_ptr;   /* [2P+1] */
  const int32_t* mol_eptr;  /* [2P+1] */
  const uint16_t* atom_w;   /* [N]    */
  const uint32_t* edge_w;   /* [Eu]   */
} imp_compact_graph_t;
int imp_mpnn_forward_fused_compact(const imp_compact_graph_t* cg, const float* d_atom_emb, int32_t atom_vocab,
                                   const float* d_bond_emb, int32_t d, int32_t bond_dim, int32_t steps, const void* d_packed,
                                   float eps, int32_t flags, int32_t max_mol_atoms, float* d_pooled, int32_t* d_status,
                                   void* stream);

/* Planned fused forward (fifth generation, the default of MPNNModel): same mathematics and the same replaced reference code
 * as imp_mpnn_forward_fused, fed from a TILE PLAN instead of the CSR arrays.
 *   imp_fused_plan     once per batch (integer-only, ~3 % of a forward): cuts the batch into self-contained 128-row tiles of
 *                      whole molecules -- best-fit over windows of 256 molecules (tiles ~97 % full instead of ~88 %), rows
 *                      listed by in-degree, CSR entries translated to tile rows -- one 2 KiB record per tile
 *                      (csrc/fused_plan.cuh).  Pass exactly one of g / cg (the int32 CSR batch or the compact feed).
 *                      This is the packed-batch counterpart of the reference's padding step (train_viscosity.py:52-59,
 *                      76-110, 291-314).  d_plan: imp_fused_plan_bytes() bytes; int32 word 4 of the buffer is a status
 *                      (0 = ok; 1 = a molecule outside the envelope: > 128 atoms, a row with > 31 entries or > 336 entries
 *                      per molecule; 2 = capacity exceeded) that the caller reads after synchronising -- such batches go
 *                      through imp_mpnn_forward_fused or the staged kernels.
 *   imp_fused_pack_planned   once per weight update and per (tower, step) -> imp_fused_pack_planned_bytes() bytes; d_packed
 *                      of the planned forward is [2 towers][steps] such blocks, cation first.
 *   imp_mpnn_forward_fused_planned   the forward (sixth generation, csrc/fused_fwd6.cu): per tile one TMA bulk copy of its
 *                      record (double-buffered); per step the packed-HFMA2 Z build and GEMM1 of the third generation, then
 *                      the gate and candidate GEMMs read the aggregated messages IN PLACE from their fp32 accumulator as a
 *                      kind::tf32 operand (no read-back / re-write of agg, one tcgen05 round trip and one context barrier
 *                      less per step).  flags: IMP_TC_FP16 [| IMP_TC_PRECISE_EPILOGUE]; IMP_TC_GEN5 selects the fifth
 *                      generation (same plan, third-generation step pipeline, weights from imp_fused_pack with
 *                      IMP_TC_FP16) for comparison.  Results do not depend on how the plan cut the batch (row order inside
 *                      a tile does not enter the arithmetic); they agree with imp_mpnn_forward_fused to a few operand ulps.
 *                      The default (sixth-generation) kernel hands the tiles out through two ticket counters in int32 words
 *                      5 and 6 of the plan buffer, which the call zeroes: one forward at a time per plan buffer. */
int64_t imp_fused_pack_planned_bytes(int32_t d, int32_t bond_dim);
int imp_fused_pack_planned(const float* d_bond_transform /* [K,d,d] */, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                           void* d_packed, void* stream);
/* The same block with the K order of the Wc halves that the seventh generation's Z fragments use (flag IMP_TC_GEN7). */
int imp_fused_pack_planned7(const float* d_bond_transform /* [K,d,d] */, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                            void* d_packed, void* stream);
int64_t imp_fused_plan_bytes(int32_t n_pairs, int32_t n_atoms, int32_t n_unique, int32_t max_mol_atoms);
int imp_fused_plan(const imp_graph_t* g, const imp_compact_graph_t* cg, int32_t atom_vocab, int32_t max_mol_atoms, void* d_plan,
                   int64_t plan_bytes, void* stream);
/* The plan from the NARROW compact feed: cg->edge_w points to 16-bit entry words
 *   src (index inside its molecule, < 128) | bond << 7 (< 256) | (multiplicity - 1) << 15   (multiplicity 1 or 2)
 * -- 0.30 instead of 0.49 KB per pair over PCIe for a streamed sweep (MPNNModel.predict_stream picks it when the batch fits). */
int imp_fused_plan_compact16(const imp_compact_graph_t* cg, int32_t atom_vocab, int32_t max_mol_atoms, void* d_plan,
                             int64_t plan_bytes, void* stream);
int imp_mpnn_forward_fused_planned(const void* d_plan, int32_t n_pairs, int32_t n_atoms, int32_t n_cat_atoms,
                                   int32_t bond_vocab, const float* d_atom_emb, int32_t atom_vocab, const float* d_bond_emb,
                                   int32_t d, int32_t bond_dim, int32_t steps, const void* d_packed, float eps, int32_t flags,
                                   float* d_pooled, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Wide atom states on the tensor cores (atom_dim 256, bond_dim 8: BASELINE configs[4], the "wide/deep" variant).
 * Replaces the same reference code as the fused forward (train_viscosity.py:163,171-187 + models/layers.py:57-164)
 * at a width where the layers are tensor-bound: per step three pipelined tcgen05 GEMM kernels (TMA bulk-copy
 * operand stages, fp32 accumulators in TMEM, IEEE-half operands) --
 *   messages      agg = Z . Wc with Z (K = 2048) built on the fly from the CSR by the producer warps,
 *   gates         z, r*h = sigma([h|agg] [Wz|Wr] + b) written as 16-bit operands,
 *   candidate     tanh([r*h|agg] Wh + bh), blend, LayerNorm, residual in the epilogue (the row lives in TMEM) --
 * between an Embedding kernel and a GlobalSumPool kernel.  Any molecule size, any number of steps.
 *   imp_wide_pack        once per weight update and per (tower, step) -> imp_wide_pack_bytes() bytes;
 *                        d_packed of the forward is [2 towers][steps] such blocks, cation first.
 *   d_workspace          imp_wide_workspace_bytes(n_atoms, d) bytes (fp32 state + four 16-bit operand matrices,
 *                        tile-packed layouts, csrc/wide_tc.cu).
 *   flags                IMP_TC_FP16 is required; IMP_TC_PRECISE_EPILOGUE as above.
 *   d_pooled             [2 * n_pairs, d] molecule sums -> imp_readout_visc / imp_readout_mp.
 * ------------------------------------------------------------------------------------------- */
int64_t imp_wide_pack_bytes(int32_t d, int32_t bond_dim);
int imp_wide_pack(const float* d_bond_transform /* [K,d,d] */, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                  void* d_packed, void* stream);
int64_t imp_wide_workspace_bytes(int32_t n_atoms, int32_t d);
int imp_mpnn_forward_wide(const imp_graph_t* g, const float* d_atom_emb, int32_t atom_vocab, const float* d_bond_emb,
                          int32_t d, int32_t bond_dim, int32_t steps, const void* d_packed, float eps, int32_t flags,
                          void* d_workspace, float* d_pooled, void* stream);
/* The stages of imp_mpnn_forward_wide as separate calls on the same workspace (layer granularity: Embedding;
 * BondMatrixMessage o Reduce; GatedUpdate gates; GatedUpdate candidate + LayerNorm + residual; GlobalSumPool). */
int imp_wide_embed(const imp_graph_t* g, const float* d_atom_emb, int32_t atom_vocab, int32_t d, void* d_workspace, void* stream);
int imp_wide_message(const imp_graph_t* g, const float* d_bond_emb, int32_t d, int32_t bond_dim, const void* d_packed_cat,
                     const void* d_packed_an, int32_t flags, void* d_workspace, void* stream);
int imp_wide_gates(const imp_graph_t* g, int32_t d, const void* d_packed_cat, const void* d_packed_an, int32_t flags,
                   void* d_workspace, void* stream);
int imp_wide_candidate(const imp_graph_t* g, int32_t d, const void* d_packed_cat, const void* d_packed_an, float eps,
                       int32_t flags, void* d_workspace, void* stream);
/* GatedUpdate as one kernel (gates stay in TMEM / shared memory); imp_wide_gates + imp_wide_candidate are the two-kernel
 * form of the same layer (flag IMP_TC_WIDE_SPLIT_GRU of imp_mpnn_forward_wide), kept for comparison. */
int imp_wide_gated_update(const imp_graph_t* g, int32_t d, const void* d_packed_cat, const void* d_packed_an, float eps,
                          int32_t flags, void* d_workspace, void* stream);
int imp_wide_pool(const imp_graph_t* g, int32_t d, const void* d_workspace, float* d_pooled, void* stream);

/* keras.layers.Dense on [rows, in_dim] fp32 device rows: y = act(x . kernel + bias); kernel (in_dim, out_dim) row-major,
 * activation 0 = linear, 1 = relu (train_viscosity.py:189,197-198,204).  Layer-level form for code written against
 * models/layers.py-style graphs; the model path uses the fused imp_pool_head_* / imp_readout_* kernels. */
int imp_dense(const float* d_x, int64_t rows, int32_t in_dim, int32_t out_dim, const float* d_kernel, const float* d_bias,
              int32_t activation, float* d_y, void* stream);

/* K6 without the pooling stage: Dense(fp, relu), Dense(mix, relu) per tower, AddTwoTensors, head
 * (train_viscosity.py:189-214 / train_melting_point.py:173-198) on molecule sums [2P, d] (cations first). */
int imp_readout_visc(const float* d_pooled, int32_t n_pairs, int32_t d, int32_t fp, int32_t mix,
                     const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an, const float* d_W_head,
                     const float* d_b_head, const float* d_T, float* d_out, float* d_aux, void* stream);
int imp_readout_mp(const float* d_pooled, int32_t n_pairs, int32_t d, int32_t fp, int32_t mix, int32_t fp2,
                   const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an, const float* d_W1,
                   const float* d_b1, const float* d_W2, const float* d_b2, float* d_out, float* d_aux, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step (train_viscosity.py:227-230,328-338: mse + l2 kernel regularisers, Adam(1e-3, clipnorm=1.0)).
 * fp32 backward kernels for atom_dim 32; every reduction is two-stage in a fixed order (bit-reproducible).
 * The forward that precedes them is the staged fp32 path with h_0..h_S and agg_0..agg_{S-1} kept.
 * "workspace" sizes come from the imp_*_workspace_floats queries; gradients of one GatedUpdate layer are laid out
 * [dWz (2d,d) | dbz | dWr | dbr | dWh | dbh | dgamma | dbeta] (the Keras variable order of the layer); gradients of the readout
 * [per tower: dW_fp, db_fp, dW_mix, db_mix | dW_head (mix,3 or mix,fp2), db_head | (mp) dW2, db2].
 * ------------------------------------------------------------------------------------------- */
/* K2 for training: lane-interleaved tables and their transposes (for imp_message_agg_bwd). */
int imp_bond_table_train(const float* d_bond_emb, int32_t bond_vocab, int32_t bond_dim, int32_t d, int32_t n_tables,
                         const float* const* h_W, float* const* h_table_il, float* const* h_table_ilT, void* stream);
/* B6: loss and readout backward.  scale = 2 / global batch (d mse / d out); *d_sse = sum of squared errors of this
 * batch; d_dpooled [2P,d]; d_out optional predictions. */
int64_t imp_readout_bwd_workspace_floats(int32_t d, int32_t fp, int32_t mix, int32_t fp2);
int imp_readout_bwd(const float* d_pooled, int32_t n_pairs, int32_t d, int32_t fp, int32_t mix, int32_t fp2,
                    const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an, const float* d_W1,
                    const float* d_b1, const float* d_W2, const float* d_b2, const float* d_T, const float* d_y, float scale,
                    float* d_dpooled, float* d_out, float* d_grads, float* d_sse, float* d_workspace, void* stream);
/* B5: GlobalSumPool backward (models/layers.py:161-164): dh[v] = [atom_id[v] > 0] * d_pooled[molecule of v]. */
int imp_pool_bwd(const int32_t* d_mol_ptr, const int32_t* d_atom_id, int32_t n_mols, const float* d_dpooled, int32_t d,
                 float* d_dh, void* stream);
/* B4: GatedUpdate backward (models/layers.py:142-156); recomputes the gates from (h, agg). */
int64_t imp_gated_update_bwd_workspace_floats(int32_t d);
/* Training form of the pair: imp_gated_update_train is imp_gated_update that also keeps the gates z, r and the candidate
 * tanh(.) of every atom ([N, d] each); imp_gated_update_bwd_stored reads them instead of recomputing the three Dense layers
 * (a third of the backward kernel's arithmetic) and returns the same gradients as imp_gated_update_bwd. */
int imp_gated_update_train(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                           const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out, float* d_z,
                           float* d_r, float* d_ht, void* stream);
/* imp_gated_update / imp_gated_update_train on the tensor cores with fp32-class accuracy (csrc/fwd_tc32.cu): both Dense
 * products of GatedUpdate.call (models/layers.py:146-151) as tcgen05.mma kind::tf32 with every operand split into two tf32
 * terms (3 MMAs per product), fp32 epilogue (expf / tanhf / sqrtf).  d_z / d_r / d_ht: all three (training form: truncation
 * splits) or all NULL (inference form: both terms rounded to tf32). */
int imp_gated_update_tc32(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                          const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out, float* d_z,
                          float* d_r, float* d_ht, void* stream);
/* The same contract on the tensor cores (csrc/bwd_tc.cu): the six contractions of the GatedUpdate backward as tcgen05.mma
 * kind::tf32 with every operand split into two tf32 terms (3 MMAs per product: fp32-class accuracy, the 2e-4 gradient bound
 * of the fp32 kernels holds); the weight-gradient accumulator stays in tensor memory for all tiles of a CTA. */
int imp_gated_update_bwd_tc(const float* d_h, const float* d_agg, const float* d_z, const float* d_r, const float* d_ht,
                            const float* d_gout, int32_t n_atoms, int32_t n_cat_atoms, int32_t d, const imp_gru_weights_t* w_cat,
                            const imp_gru_weights_t* w_an, float eps, float* d_dh, float* d_dagg, float* d_grads_cat,
                            float* d_grads_an, float* d_workspace, void* stream);
int imp_gated_update_bwd_stored(const float* d_h, const float* d_agg, const float* d_z, const float* d_r, const float* d_ht,
                                const float* d_gout, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_dh, float* d_dagg,
                                float* d_grads_cat, float* d_grads_an, float* d_workspace, void* stream);
int imp_gated_update_bwd(const float* d_h, const float* d_agg, const float* d_gout, int32_t n_atoms, int32_t n_cat_atoms,
                         int32_t d, const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_dh,
                         float* d_dagg, float* d_grads_cat, float* d_grads_an, float* d_workspace, void* stream);
/* B3a: dh += sum_e mult_e T[b_e]^T dagg[src_e] over the rows of the SAME CSR (the live edge set is symmetric:
 * train_viscosity.py:87-91 appends the reverse of every entry).  Tables from imp_bond_table_train. */
int imp_message_agg_bwd(const imp_graph_t* g, const float* d_dagg, int32_t d, const float* d_table_ilT_cat,
                        const float* d_table_ilT_an, float* d_dh /* accumulated */, void* stream);
/* B3b: gradients of bond_transform (both towers of one step) and of the shared bond embedding (accumulated).
 * d_entry_dst[e] = destination atom of CSR entry e; chunks = host-built split of the (tower, bond) buckets:
 * chunk c covers bucket_perm[chunk_begin[c] .. chunk_end[c]); bucket_chunk_ptr[2*V_b+1] delimits each bucket's chunks. */
int imp_bond_transform_bwd(const imp_graph_t* g, const int32_t* d_entry_dst, const int32_t* d_chunk_begin,
                           const int32_t* d_chunk_end, int32_t n_chunks, const int32_t* d_bucket_chunk_ptr,
                           const float* d_dagg, const float* d_h, int32_t d, int32_t bond_dim, const float* d_bond_emb,
                           const float* d_W_cat, const float* d_W_an, float* d_dW_cat, float* d_dW_an, float* d_dbond_emb,
                           float* d_dtable, float* d_workspace, void* stream);
/* B1: Embedding(atom) backward. */
int64_t imp_embed_bwd_workspace_floats(int32_t atom_vocab, int32_t d);
int imp_embed_bwd(const int32_t* d_atom_id, const float* d_dh0, int32_t n_atoms, int32_t atom_vocab, int32_t d,
                  float* d_datom_emb, float* d_workspace, void* stream);
/* Per-variable clip_by_norm + Adam over a flat parameter buffer [Keras semantics].  var_off[n_vars+1] delimits the
 * variables; var_l2[v] = l2 regulariser coefficient of variable v (its gradient 2*l2*w is added before clipping). */
int imp_clip_adam(float* d_param, const float* d_grad, float* d_m, float* d_v, const int64_t* d_var_off,
                  const float* d_var_l2, int32_t n_vars, float* d_norms2, float clipnorm, float lr, float beta1, float beta2,
                  float eps, int32_t step, void* stream);
/* The same step with the reference's exact Embedding semantics and a distributed mean [Keras semantics, TF / Keras 2.12;
 * oracle/ref_model.py:adam_step]: the optimizer clips an IndexedSlices gradient BEFORE de-duplicating it, so the clip norm
 * of the two Embedding variables (train_viscosity.py:163-164) is taken over the per-occurrence rows.  The n_occ variables
 * d_occ_var[i] (device int32) use d_occ_norm2[i] (device; from imp_sumsq / imp_bond_occurrence_norm2) as their squared clip
 * norm.  d_pair_count (device scalar, optional): the gradients and occurrence norms were computed for the SUM over pairs
 * and are turned into the mean by 1 / *d_pair_count -- the count rides in the same all-reduce as the gradient bucket.
 * d_norms2 holds 2 * n_vars floats (squared norms, then the l2 loss terms).  d_loss (optional, needs d_sse = sum of squared
 * errors): receives this step's loss = sse / count (or sse * inv_batch when d_pair_count is NULL) + the l2 terms, computed
 * with the weights BEFORE the update, as Keras reports it (train_viscosity.py:189,227-230). */
int imp_clip_adam_sparse(float* d_param, const float* d_grad, float* d_m, float* d_v, const int64_t* d_var_off,
                         const float* d_var_l2, int32_t n_vars, float* d_norms2, float clipnorm, float lr, float beta1,
                         float beta2, float eps, int32_t step, int32_t n_occ, const int32_t* d_occ_var,
                         const float* d_occ_norm2, const float* d_pair_count, const float* d_sse, float inv_batch,
                         float* d_loss, void* stream);
/* Keras `evaluate` loss on device predictions: mean((pred - y)^2) + sum_v var_l2[v] * |param_v|^2, fixed summation order.
 * d_scratch: 1 + n_vars floats. */
int imp_eval_loss(const float* d_pred, const float* d_y, int64_t n, const float* d_param, const int64_t* d_var_off,
                  const float* d_var_l2, int32_t n_vars, float* d_scratch, float* d_loss, void* stream);
/* Deterministic sum of squares of n floats (two-stage, fixed order); workspace >= 1024 floats.  With x = dL/dh0 [N, d]
 * (output of the backward pass before imp_embed_bwd) this is the per-occurrence squared norm of the atom Embedding. */
int imp_sumsq(const float* d_x, int64_t n, float* d_out, float* d_workspace, void* stream);
/* Per-occurrence squared gradient norm of the bond Embedding (csrc/occ_norm.cu): sum over CSR entries e of
 * mult_e * |g_e|^2, g_e[k] = sum_s dagg_s[dst_e]^T W_{s,k} h_s[src_e] (models/layers.py:108-112 under autodiff).
 * d_h_steps / d_dagg_steps: HOST arrays of `steps` device pointers ([N, d] fp32: the state entering step s, the gradient of
 * its aggregated messages); d_packed_*: `steps` images of imp_occ_pack per tower; n_cat_unique = row_ptr[n_cat_atoms];
 * workspace: imp_bond_occurrence_norm2_workspace_floats(n_unique) floats.  atom_dim 32, bond_dim 8 (tcgen05). */
int64_t imp_occ_pack_bytes(int32_t d, int32_t bond_dim);
int imp_occ_pack(const float* d_bond_transform, int32_t d, int32_t bond_dim, void* d_packed, void* stream);
int64_t imp_bond_occurrence_norm2_workspace_floats(int32_t n_unique);
int imp_bond_occurrence_norm2(const imp_graph_t* g, int32_t n_cat_unique, const int32_t* d_entry_dst, int32_t steps,
                              const float* const* d_h_steps, const float* const* d_dagg_steps, int32_t d, int32_t bond_dim,
                              const void* d_packed_cat, const void* d_packed_an, float* d_out, float* d_workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Diagnostics: one-CTA tcgen05 product D[128,N] = A[128,K] * B[N,K]^T (kind 0 = bf16, 1 = tf32 operands from
 * shared memory; 2 = bf16, 3 = f16 with the A operand written to tensor memory by the threads, the form the fused
 * forward uses; fp32 accumulate) through the library's own shared-memory staging layout and UMMA descriptors.  Used by
 * tests/test_gpu_tensor.py to validate the tensor-core plumbing in isolation.
 * ------------------------------------------------------------------------------------------- */
int imp_tc_selftest(const float* d_A, const float* d_B, float* d_D, int32_t N, int32_t K, int32_t kind,
                    int32_t swap_lbo_sbo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Transfer-learning head (train_melting_point_transfer.py:76-106, 189-241): the viscosity model cut at "mix_cat_an" +
 * Dense(256, relu) -> BatchNormalization -> Dense(128, relu) -> Dropout(0.3) -> Dense(64, relu) -> Dense(1), Huber(delta = 1),
 * Adam without clipping, layers frozen / unfrozen by name.  Host sequence: ionic_mpnn_b200/transfer.py.  One row per ion
 * pair; every reduction walks the rows in order (bit-reproducible).
 *   imp_dense_bwd      gradients of imp_dense: d_y = the layer's output (relu mask; may be NULL for activation 0);
 *                      d_gx [rows, in] (optional), d_gkernel [in, out] + d_gbias [out] (optional, need d_x).
 *   imp_batchnorm      keras BatchNormalization on [rows, channels] (non-fused path: biased batch variance; moving averages
 *                      updated in place when training != 0; inference uses them).  d_save_mean / d_save_inv [channels]: the
 *                      statistics the backward pass needs (both or neither).
 *   imp_dropout        y = x * keep / (1 - rate) with keep(i) a pure function of (seed, i): applying it to the upstream
 *                      gradient with the same seed is the backward pass.
 *   imp_huber          *d_loss_sum = sum_i huber(pred_i - target_i); d_dpred_i = scale * clip(pred_i - target_i, -delta, delta).
 *   imp_add            y = a + b (AddTwoTensors, models/layers.py:44-49, as its own call for the layer API).
 * ------------------------------------------------------------------------------------------- */
int imp_dense_bwd(const float* d_x, const float* d_y, const float* d_gy, int64_t rows, int32_t in_dim, int32_t out_dim,
                  const float* d_kernel, int32_t activation, float* d_gx, float* d_gkernel, float* d_gbias, void* stream);
int imp_batchnorm(const float* d_x, int64_t rows, int32_t channels, const float* d_gamma, const float* d_beta,
                  float* d_moving_mean, float* d_moving_var, float momentum, float eps, int32_t training, float* d_y,
                  float* d_save_mean, float* d_save_inv, void* stream);
int imp_batchnorm_bwd(const float* d_x, const float* d_gy, int64_t rows, int32_t channels, const float* d_gamma,
                      const float* d_save_mean, const float* d_save_inv, float* d_gx, float* d_ggamma, float* d_gbeta, void* stream);
int imp_dropout(const float* d_x, int64_t n, float rate, uint64_t seed, float* d_y, void* stream);
int imp_huber(const float* d_pred, const float* d_target, int64_t n, float delta, float scale, float* d_loss_sum, float* d_dpred,
              void* stream);
int imp_add(const float* d_a, const float* d_b, int64_t n, float* d_y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IMP_B200_H */
