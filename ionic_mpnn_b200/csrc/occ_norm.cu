// Per-occurrence gradient norms of the two Embedding variables -- what Keras' clipnorm sees for them.
//
// Reference: train_viscosity.py:163-164 (shared `Embedding` layers), :227-230 (`Adam(1e-3, clipnorm=1.0)`).
// [Keras semantics, TF / Keras 2.12] the gradient of an Embedding variable is an IndexedSlices whose rows are the gradients
// with respect to every looked-up row (one per (tower, sample, atom slot) / (tower, sample, edge slot)); the optimizer clips
// BEFORE it de-duplicates, and tf.clip_by_norm takes the norm over those un-deduplicated rows:
//     norm^2(atom_emb) = sum over atoms n           |dL/dh0[n]|^2
//     norm^2(bond_emb) = sum over edge entries e    |g_e|^2,   g_e[k] = sum_s dagg_s[dst_e]^T W_{s,k} h_s[src_e]
// (oracle/ref_model.py:adam_step restates this; padded and masked slots have zero gradient and do not contribute).
// The packed CSR keeps the reference's duplicate entries as a multiplicity: such an entry stands for `mult` occurrences with
// the same g_e, i.e. it contributes mult * |g_e|^2.
//
// imp_sumsq: deterministic two-stage sum of squares (the atom part: x = dL/dh0, [N, d]).
// imp_bond_occurrence_norm2: one persistent tcgen05 kernel over tiles of 128 consecutive CSR entries of one tower.  Per step s
// the source rows h_s[src_e] are gathered into the shared-memory A operand (fp32 -> IEEE half, 8 lanes per 128-byte row),
// Y = H . Wcat_s^T (M 128, N 256 = (k, l), K 32 = m; Wcat_s[k*32+l][m] = W_s[k][l][m]) lands in 256 TMEM columns, and the
// thread that owns entry e contracts its TMEM lane with the fp32 row dagg_s[dst_e]: 8 dot products of 32 terms, accumulated
// over the steps in registers.  mult_e * |g_e|^2 is reduced per tile in a fixed order, the tile partials by a second kernel:
// the result is bit-reproducible.  The 16-bit operands perturb each g_e by ~5e-4 relative, unbiased: the norm over ~10^2..10^7
// entries is accurate to ~1e-5, and only matters when it exceeds clipnorm.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

extern "C" int imp_device_is_sm100(void);

namespace imp {
namespace occ {

constexpr int D = 32, KB = 8, TILE = 128, NCOL = KB * D;  // 256 accumulator columns
constexpr int A_BYTES = TILE * D * 2;                     // 8 KB
constexpr int B_BYTES = NCOL * D * 2;                     // 16 KB per (tower, step)
constexpr int MAX_STEPS = 8;
constexpr int STG_LD = D + 1;
constexpr int RED_BLOCKS = 1024;

// ---- deterministic sum of squares: RED_BLOCKS partials (grid-stride, fixed trees), then one block
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ part) {
  __shared__ float red[256];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s = fmaf(x[i], x[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// Wcat image: element (n = k * 32 + l, kk = m) = W[k][l][m], canonical K-major SWIZZLE_NONE layout with R = 256 rows
__global__ void occ_pack_kernel(const float* __restrict__ W, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NCOL * D) return;
  const int n = i / D, m = i % D;
  *reinterpret_cast<uint16_t*>(out + tc::chunk_off(n, m / 8, NCOL) + (m % 8) * 2) = tc::cvt16<tc::FMT_F16>(W[n * D + m]);
}

struct OccArgs {
  const float* h[MAX_STEPS];
  const float* dagg[MAX_STEPS];
  const int32_t* col_src;
  const int32_t* edge_bm;
  const int32_t* entry_dst;
  const uint8_t* packed_cat;  // [steps][B_BYTES]
  const uint8_t* packed_an;
  float* part;                // [tiles_cat + tiles_an]
  int steps, e_split, n_unique, tiles_cat, tiles_an, n_cta_cat;
};

__global__ void __launch_bounds__(TILE) bond_occ_norm_kernel(const OccArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;                                        // steps * B_BYTES
  uint8_t* sA = smem + a.steps * B_BYTES;                    // A_BYTES
  float* stg = reinterpret_cast<float*>(sA + A_BYTES);       // TILE * STG_LD floats: dagg rows of the tile
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[4];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int tower = (int)blockIdx.x >= a.n_cta_cat;
  const int n_cta = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const int n_tiles = tower ? a.tiles_an : a.tiles_cat;
  const int e_begin = tower ? a.e_split : 0, e_end = tower ? a.n_unique : a.e_split;
  {
    const uint4* src = reinterpret_cast<const uint4*>(tower ? a.packed_an : a.packed_cat);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    for (int i = t; i < a.steps * (B_BYTES / 16); i += TILE) dst[i] = __ldg(src + i);
  }
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<NCOL>(&tmem_slot);
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc = tc::make_idesc(tc::FMT_F16, TILE, NCOL);
  const uint64_t dA = tc::make_smem_desc(tc::smem_u32(sA), TILE * 16, 128);
  const uint64_t dB0 = tc::make_smem_desc(tc::smem_u32(sB), NCOL * 16, 128);
  const int g = lane >> 3, q = lane & 7;
  uint32_t ph = 0;

  for (int tile = cta; tile < n_tiles; tile += n_cta) {
    const int e = e_begin + tile * TILE + t;
    const bool valid = e < e_end;
    int src = 0, dst = 0;
    float mult = 0.f;
    if (valid) {
      src = __ldg(a.col_src + e);
      dst = __ldg(a.entry_dst + e);
      mult = (float)((uint32_t)__ldg(a.edge_bm + e) >> 16);
    }
    float acc[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) acc[k] = 0.f;
    for (int s = 0; s < a.steps; ++s) {
      const float* hs = a.h[s];
      const float* ds = a.dagg[s];
      // 8 lanes per 128-byte row: h_s[src] -> A operand (half), dagg_s[dst] -> padded fp32 staging tile
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = 4 * it + g;
        const int rs = __shfl_sync(0xffffffffu, src, r), rd = __shfl_sync(0xffffffffu, dst, r);
        const bool rv = __shfl_sync(0xffffffffu, (int)valid, r) != 0;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
        if (rv) {
          x = __ldg(reinterpret_cast<const float4*>(hs + (int64_t)rs * D) + q);
          y = __ldg(reinterpret_cast<const float4*>(ds + (int64_t)rd * D) + q);
        }
        *reinterpret_cast<uint2*>(sA + (q >> 1) * (TILE * 16) + (warp * 32 + r) * 16 + (q & 1) * 8) =
            make_uint2(tc::pack_f16x2(x.x, x.y), tc::pack_f16x2(x.z, x.w));
        float* sr = stg + (warp * 32 + r) * STG_LD + 4 * q;
        sr[0] = y.x, sr[1] = y.y, sr[2] = y.z, sr[3] = y.w;
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      __syncthreads();
      if (warp == 0) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
          const uint64_t dB = dB0 + (uint64_t)(s * (B_BYTES / 16));
          tc::mma_bf16(tmem, dA, dB, idesc, false);
          tc::mma_bf16(tmem, dA + (uint64_t)((2 * TILE * 16) >> 4), dB + (uint64_t)((2 * NCOL * 16) >> 4), idesc, true);
          tc::mma_commit(&bar);
        }
        __syncwarp();
      }
      float dg[D];
#pragma unroll
      for (int c = 0; c < D; ++c) dg[c] = stg[t * STG_LD + c];
      tc::mbar_wait(&bar, ph);
      ph ^= 1;
      tc::fence_after_thread_sync();
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        float v[32];
        tc::tmem_ld32(tmem + lane_off + (uint32_t)(k * D), v);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < D; c += 4) {
          s0 = fmaf(dg[c], v[c], s0), s1 = fmaf(dg[c + 1], v[c + 1], s1);
          s2 = fmaf(dg[c + 2], v[c + 2], s2), s3 = fmaf(dg[c + 3], v[c + 3], s3);
        }
        acc[k] += (s0 + s1) + (s2 + s3);
      }
      tc::fence_before_thread_sync();
      __syncthreads();  // operands, staging tile and accumulator columns are free for the next step
      tc::fence_after_thread_sync();
    }
    float val = 0.f;
#pragma unroll
    for (int k = 0; k < KB; ++k) val = fmaf(acc[k], acc[k], val);
    val = valid ? val * mult : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if (lane == 0) red[warp] = val;
    __syncthreads();
    if (t == 0) a.part[(tower ? a.tiles_cat : 0) + tile] = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc<NCOL>(tmem);
  }
}

}  // namespace occ
}  // namespace imp

using namespace imp;

extern "C" int imp_sumsq(const float* d_x, int64_t n, float* d_out, float* d_workspace, void* stream) {
  IMP_REQUIRE(n >= 0 && d_out && d_workspace && (n == 0 || d_x), IMP_ERR_ARG, "imp_sumsq: bad arguments");
  occ::sumsq_partial_kernel<<<occ::RED_BLOCKS, 256, 0, (cudaStream_t)stream>>>(d_x, n, d_workspace);
  IMP_LAUNCH_CHECK();
  occ::sum_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_workspace, occ::RED_BLOCKS, d_out);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t imp_occ_pack_bytes(int32_t d, int32_t bond_dim) {
  return (d == occ::D && bond_dim == occ::KB) ? (int64_t)occ::B_BYTES : (int64_t)IMP_ERR_DIM;
}

extern "C" int imp_occ_pack(const float* d_bond_transform, int32_t d, int32_t bond_dim, void* d_packed, void* stream) {
  IMP_REQUIRE(d_bond_transform && d_packed, IMP_ERR_ARG, "imp_occ_pack: null pointer");
  IMP_REQUIRE(d == occ::D && bond_dim == occ::KB, IMP_ERR_DIM, "imp_occ_pack: built for atom_dim %d, bond_dim %d (got %d, %d)", occ::D,
              occ::KB, d, bond_dim);
  occ::occ_pack_kernel<<<(occ::NCOL * occ::D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_bond_transform, (uint8_t*)d_packed);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t imp_bond_occurrence_norm2_workspace_floats(int32_t n_unique) {
  return ceil_div(n_unique > 0 ? n_unique : 0, occ::TILE) + 2;
}

extern "C" int imp_bond_occurrence_norm2(const imp_graph_t* g, int32_t n_cat_unique, const int32_t* d_entry_dst, int32_t steps,
                                         const float* const* d_h_steps,
                                         const float* const* d_dagg_steps, int32_t d, int32_t bond_dim, const void* d_packed_cat,
                                         const void* d_packed_an, float* d_out, float* d_workspace, void* stream) {
  IMP_REQUIRE(g && d_out && d_workspace && d_h_steps && d_dagg_steps, IMP_ERR_ARG, "imp_bond_occurrence_norm2: null pointer");
  IMP_REQUIRE(d == occ::D && bond_dim == occ::KB, IMP_ERR_DIM, "imp_bond_occurrence_norm2: built for atom_dim %d, bond_dim %d (got %d, %d)",
              occ::D, occ::KB, d, bond_dim);
  IMP_REQUIRE(steps >= 1 && steps <= occ::MAX_STEPS, IMP_ERR_DIM, "imp_bond_occurrence_norm2: 1..%d steps (got %d)", occ::MAX_STEPS, steps);
  cudaStream_t st = (cudaStream_t)stream;
  if (g->n_unique <= 0) {
    IMP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), st));
    return 0;
  }
  IMP_REQUIRE(g->col_src && g->edge_bm && d_entry_dst && d_packed_cat && d_packed_an, IMP_ERR_ARG,
              "imp_bond_occurrence_norm2: null pointer");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_bond_occurrence_norm2: tcgen05 needs an sm_100 device");
  occ::OccArgs a{};
  for (int s = 0; s < steps; ++s) {
    IMP_REQUIRE(d_h_steps[s] && d_dagg_steps[s], IMP_ERR_ARG, "imp_bond_occurrence_norm2: null step pointer");
    a.h[s] = d_h_steps[s], a.dagg[s] = d_dagg_steps[s];
  }
  // entries of the cation tower are the CSR rows [0, n_cat_atoms): n_cat_unique = row_ptr[n_cat_atoms], known to the host
  IMP_REQUIRE(n_cat_unique >= 0 && n_cat_unique <= g->n_unique, IMP_ERR_ARG, "imp_bond_occurrence_norm2: bad n_cat_unique");
  const int32_t e_split = n_cat_unique;
  a.col_src = g->col_src, a.edge_bm = g->edge_bm, a.entry_dst = d_entry_dst;
  a.packed_cat = (const uint8_t*)d_packed_cat, a.packed_an = (const uint8_t*)d_packed_an;
  a.part = d_workspace;
  a.steps = steps, a.e_split = e_split, a.n_unique = g->n_unique;
  a.tiles_cat = (int)ceil_div(e_split, occ::TILE), a.tiles_an = (int)ceil_div(g->n_unique - e_split, occ::TILE);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int slots = 2 * sms;  // 256 TMEM columns per CTA: two CTAs per SM
  const int total = a.tiles_cat + a.tiles_an;
  int n_cat = (int)((int64_t)slots * a.tiles_cat / (total > 0 ? total : 1));
  if (a.tiles_cat > 0 && n_cat < 1) n_cat = 1;
  if (n_cat > a.tiles_cat) n_cat = a.tiles_cat;
  int n_an = slots - n_cat;
  if (n_an > a.tiles_an) n_an = a.tiles_an;
  a.n_cta_cat = n_cat;
  const size_t smem = (size_t)steps * occ::B_BYTES + occ::A_BYTES + occ::TILE * occ::STG_LD * 4;
  IMP_CUDA(cudaFuncSetAttribute(occ::bond_occ_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  occ::bond_occ_norm_kernel<<<n_cat + n_an, occ::TILE, smem, st>>>(a);
  IMP_LAUNCH_CHECK();
  occ::sum_partials_kernel<<<1, 256, 0, st>>>(d_workspace, total, d_out);
  IMP_LAUNCH_CHECK();
  return 0;
}
