// Shared pieces of the fused whole-tower forward kernels (fused_fwd.cu: generations 1-3, fused_fwd4.cu: generation 4).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace imp {


constexpr int FZ_ROWS = 128;      // rows per tile = TMEM lanes
constexpr int FZ_D = 32;          // atom_dim
constexpr int FZ_K = 8;           // bond_dim
constexpr int FZ_GROUP = 64;      // molecules per scheduling unit
constexpr int FZ_HS = 36;         // floats per h row in shared memory (144 B: conflict-free 16-byte row reads)
constexpr int FZ_CS = 12;         // floats per bond-coefficient row (48 B)
constexpr int FZ_MAX_STEPS = 4;
constexpr int FZ_MAX_VB = 256;

struct FusedPack {  // one (tower, step)
  static constexpr int WC_BYTES = FZ_D * (FZ_D * FZ_K) * 2;  // Wc[n = l][kk = m*8+k], chunk-major, 16 KiB
  static constexpr int BZR_BYTES = 2 * FZ_D * 2 * FZ_D * 2;  // [Wz | Wr]^T, 8 KiB
  static constexpr int BH_BYTES = FZ_D * 2 * FZ_D * 2;       // Wh^T, 4 KiB
  static constexpr int BIAS_FLOATS = 5 * FZ_D;               // bz, br, bh, gamma, beta
  static constexpr int OFF_BZR = WC_BYTES;
  static constexpr int OFF_BH = OFF_BZR + BZR_BYTES;
  static constexpr int OFF_BIAS = OFF_BH + BH_BYTES;
  static constexpr int OFF_BBZR = OFF_BIAS + BIAS_FLOATS * 4;  // [64 x 16] K-major block, column 0 = 0.5 * (bz | br)
  static constexpr int BBZR_BYTES = 2 * FZ_D * 16 * 2;         //   (third-generation kernel: biases ride in the GEMM)
  static constexpr int OFF_BBH = OFF_BBZR + BBZR_BYTES;        // [32 x 16] block, column 0 = bh
  static constexpr int BBH_BYTES = FZ_D * 16 * 2;
  static constexpr int BYTES = OFF_BBH + BBH_BYTES;            // 32 384
};
static_assert(FusedPack::BYTES % 128 == 0, "pack must keep 128-byte alignment of the next step");

struct FusedArgs {
  const int* mol_ptr;
  const int* atom_id;
  const int* row_ptr;
  const int* col_src;
  const int* edge_bm;
  const float* atom_emb;
  const float* bond_emb;
  const unsigned char* packed;  // [2][steps][FusedPack::BYTES]
  float* pooled;                // [2P][32]
  int* status;                  // optional: set to 1 if a molecule does not fit one tile
  int n_pairs, atom_vocab, bond_vocab, steps, n_cta_cat;
  int n_atoms, n_unique;
  float eps;
  // compact input feed (imp_mpnn_forward_fused_compact): 16-bit atom words, 32-bit entry words, per-molecule entry offsets
  const int* mol_eptr;             // [2P+1] first CSR entry of every molecule
  const unsigned short* atom_w;    // [N] atom id | in-degree << 8
  const unsigned int* edge_w;      // [Eu] src (molecule-local) | bond << 8 | multiplicity << 16
  long long* prof;  // FZ_PROFILE builds only
  int debug;        // FZ_PROFILE builds only (timing ablations, results are wrong): 1 = skip the entry loop, 2 = skip MMAs
};

struct FusedCtl {
  uint32_t tmem_base;
  uint32_t pad;
  uint64_t wbar;  // mbarrier of the TMA weight load (third-generation kernel)
};

__device__ __forceinline__ float fz_tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool PRECISE>
__device__ __forceinline__ float fz_sigmoid(float x) {
  if (PRECISE) return 1.0f / (1.0f + expf(-x));
  return fmaf(0.5f, fz_tanh_fast(0.5f * x), 0.5f);
}
// sigmoid(2 y) for a pre-activation y that was already halved by the packed weights
template <bool PRECISE>
__device__ __forceinline__ float fz_sigmoid_half(float y) {
  if (PRECISE) return 1.0f / (1.0f + expf(-2.0f * y));
  return fmaf(0.5f, fz_tanh_fast(y), 0.5f);
}
template <bool PRECISE>
__device__ __forceinline__ float fz_tanh(float x) {
  return PRECISE ? tanhf(x) : fz_tanh_fast(x);
}

// tanh of two halves with ONE MUFU (tanh.approx.f16x2; max absolute error 2^-10.987, the same order as tanh.approx.f32)
__device__ __forceinline__ __half2 fz_tanh_h2(__half2 x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(*reinterpret_cast<const uint32_t*>(&x)));
  return *reinterpret_cast<const __half2*>(&y);
}

// rsqrt.approx.ftz: one MUFU.RSQ (rsqrtf() wraps it in a denormal rescue -- FSETP, two FMULs by 2^24 / 2^12 -- that a
// LayerNorm variance + eps >= 1e-3 never needs)
__device__ __forceinline__ float fz_rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Optional phase timing (compile with -DFZ_PROFILE): per-phase clock64 deltas of selected threads, summed into
// a.prof[thread-class][32] (thread classes: u == 0, u == 96 (warp 3), u == 224 (warp 7)); read by tools/fused_phase_profile.py.
#ifdef FZ_PROFILE
#define FZ_DEBUG(a) ((a).debug)
#define FZ_PROF_N 32
#define FZ_PROF_DECL                                                                     \
  long long prof_acc[FZ_PROF_N];                                                         \
  for (int i_ = 0; i_ < FZ_PROF_N; ++i_) prof_acc[i_] = 0;                               \
  long long prof_last = clock64();                                                       \
  const int prof_cls = (u == 0) ? 0 : (u == 96) ? 1 : (u == 224) ? 2 : -1
#define FZ_PROF_T(i)                         \
  do {                                       \
    const long long now_ = clock64();        \
    prof_acc[i] += now_ - prof_last;         \
    prof_last = now_;                        \
  } while (0)
#define FZ_PROF_FLUSH                                                                                       \
  if (prof_cls >= 0 && a.prof)                                                                               \
    for (int i_ = 0; i_ < FZ_PROF_N; ++i_) atomicAdd(reinterpret_cast<unsigned long long*>(a.prof) + prof_cls * FZ_PROF_N + i_, (unsigned long long)prof_acc[i_])
#else
#define FZ_DEBUG(a) 0
#define FZ_PROF_DECL
#define FZ_PROF_T(i)
#define FZ_PROF_FLUSH
#endif

}  // namespace imp
