// GatedUpdate forward with fp32-class accuracy on the tensor cores (tcgen05 kind::tf32, 3xTF32 operand splits), atom_dim 32.
//
// Replaces GatedUpdate.call (models/layers.py:142-156) on the 1e-5 path: the fp32 inference kernels and the forward half of
// the training step (train_viscosity.py:227-230), where the fp32 SIMT kernel (gated_update32_kernel, 12 k FMAs per atom)
// is FFMA-bound.  Same arithmetic contract: fp32 state in and out, expf / tanhf / sqrtf epilogue, biased-variance
// LayerNorm with epsilon, residual; optionally keeps the gates for the backward pass (imp_gated_update_train).
//
//   [z | r] pre-activations = [h | agg] . [Wz | Wr]      (128 x 64, K = 64)   A = the row operand in TENSOR MEMORY, hi and lo terms
//   candidate pre-activation = [r*h | agg] . Wh          (128 x 32, K = 64)   (r*h overwrites the h columns of the operand)
// every product = hi.hi + hi.lo + lo.hi with x = hi + lo, hi = x with 13 low significand bits cleared (csrc/bwd_tc.cu).
// Two threads per atom row (TMEM lane; 16 columns each); 224 TMEM columns and 52 KB of shared memory per CTA: two CTAs per SM.
#include "common.cuh"
#include "tc_common.cuh"

namespace imp {

constexpr int FT_D = 32;
constexpr int FT_TILE = 128;

struct FtSmem {
  float W1[2][64 * 64];  // B[n][k] = (n < 32 ? Wz : Wr)[k][n & 31], K-major tf32, hi then lo
  float W2[2][32 * 64];  // B[n][k] = Wh[k][n]
  float bz[FT_D], br[FT_D], bh[FT_D], gamma[FT_D], beta[FT_D];
  float xs[2][2][FT_TILE];  // LayerNorm partial sums of the two column halves of a row (sum; sum of squared deviations)
  // the next tile's h and agg rows, copied by cp.async with 8 lanes per 128-byte row (a thread loading its own row touches
  // 32 L1 lines per instruction: that, not HBM, capped the one-thread-per-row form at 2.4 TB/s); 16-byte chunk c of row r at
  // chunk c ^ (r & 7), so that the row owners read their chunks back without bank conflicts
  float in[2][FT_TILE * FT_D];
  float out[4][32 * FT_D];  // per TMEM quadrant: 32 rows on their way out (same swizzle), stored 8 lanes per 128-byte row
  uint64_t bar[2];
  uint32_t tmem_base;
};

// x = hi + lo.  RN: both terms ROUNDED to tf32 (cvt.rna): |x - hi - lo| <= 2^-22 |x|, unbiased -- the inference form (no gate
// outputs), where it separates 2.1e-5 from 1.6e-5 on the ill-conditioned predictions of tests/test_gpu_parity.py.  Otherwise the
// truncation split of csrc/bwd_tc.cu (one LOP3, lo truncated by the MMA: biased towards zero, twice as coarse, 5 % faster) --
// the training form: gradients are tested at 2e-4.
template <bool RN>
__device__ __forceinline__ void ft_split(float x, float& hi, float& lo) {
  if (RN) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
  } else {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
  }
}
__device__ __forceinline__ void ft_mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}
// expf / tanhf / IEEE division and square root, as the fp32 SIMT kernel: ex2.approx / rcp.approx forms (absolute error 3e-7, 0.1 ms
// per launch faster) moved an Adam-normalised weight of the transfer head by 2.9e-4 after three steps where the test allows
// 2e-4 (tests/test_gpu_transfer.py) -- Adam divides by sqrt(v), so rounding noise in near-zero gradients is amplified
__device__ __forceinline__ float ft_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float ft_tanh(float x) { return tanhf(x); }

// 256 threads per 128-row tile: warps q and q + 4 share TMEM quadrant q; a thread owns one row and 16 of its 32 columns (column
// half hf = warp / 4) through every phase, so a CTA has eight warps in flight instead of four and each warp's dependent chain
// (split -> store -> MMA -> activations -> MMA -> LayerNorm) is half as long.  The two LayerNorm reductions of a row cross the
// warp pair through shared memory (64-thread named barriers).  Two CTAs per SM (224 TMEM columns each): 16 warps per SM.
constexpr int FT_THREADS = 2 * FT_TILE;

template <bool RN>
__global__ void __launch_bounds__(FT_THREADS, 2) gated_update_tc32_kernel(const float* __restrict__ h, const float* __restrict__ agg,
                                                                          int n_atoms, int n_cat, int n_cta_cat, imp_gru_weights_t wc,
                                                                          imp_gru_weights_t wa, float eps, float* __restrict__ h_out,
                                                                          float* __restrict__ z_out, float* __restrict__ r_out,
                                                                          float* __restrict__ ht_out) {
  constexpr int D = FT_D, DH = FT_D / 2;
  extern __shared__ __align__(1024) unsigned char ft_raw[];
  FtSmem& s = *reinterpret_cast<FtSmem*>(ft_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;  // TMEM quadrant; column half
  const int trow = q * 32 + lane;          // row of the tile = TMEM lane
  const bool is_cat = (int)blockIdx.x < n_cta_cat;
  const imp_gru_weights_t& w = is_cat ? wc : wa;
  const int base = is_cat ? 0 : n_cat, a_end = is_cat ? n_cat : n_atoms;
  const int n_tiles = (a_end - base + FT_TILE - 1) / FT_TILE;
  const int cta = is_cat ? blockIdx.x : blockIdx.x - n_cta_cat, n_cta = is_cat ? n_cta_cat : gridDim.x - n_cta_cat;

  for (int i = tid; i < 64 * 64; i += FT_THREADS) {  // element (n, k) at chunk_off(n, k / 4, R) + (k % 4) * 4
    const int n = i / 64, k = i % 64;
    float hi, lo;
    ft_split<RN>(__ldg((n < 32 ? w.Wz : w.Wr) + k * D + (n & 31)), hi, lo);
    const int o = (tc::chunk_off(n, k / 4, 64) + (k % 4) * 4) / 4;
    s.W1[0][o] = hi, s.W1[1][o] = lo;
    if (n < 32) {
      ft_split<RN>(__ldg(w.Wh + k * D + n), hi, lo);
      const int o2 = (tc::chunk_off(n, k / 4, 32) + (k % 4) * 4) / 4;
      s.W2[0][o2] = hi, s.W2[1][o2] = lo;
    }
  }
  if (tid < D) s.bz[tid] = w.bz[tid], s.br[tid] = w.br[tid], s.bh[tid] = w.bh[tid], s.gamma[tid] = w.gamma[tid], s.beta[tid] = w.beta[tid];
  if (warp == 0) tc::tmem_alloc<256>(&s.tmem_base);
  if (tid == 0) {
    tc::mbar_init(&s.bar[0], 1);
    tc::mbar_init(&s.bar[1], 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();

  const uint32_t tm = s.tmem_base, lane_off = (uint32_t)(q * 32) << 16;
  const uint32_t tXhi = tm, tXlo = tm + 64, tD1 = tm + 128, tD2 = tm + 192;  // X = [h -> r*h | agg]
  const uint32_t id64 = tc::make_idesc(tc::FMT_TF32, FT_TILE, 64), id32 = tc::make_idesc(tc::FMT_TF32, FT_TILE, 32);
  const uint64_t dW1[2] = {tc::make_smem_desc(tc::smem_u32(s.W1[0]), 64 * 16, 128), tc::make_smem_desc(tc::smem_u32(s.W1[1]), 64 * 16, 128)};
  const uint64_t dW2[2] = {tc::make_smem_desc(tc::smem_u32(s.W2[0]), 32 * 16, 128), tc::make_smem_desc(tc::smem_u32(s.W2[1]), 32 * 16, 128)};
  const int cb = DH * hf;  // first column of this thread
  auto to_tmem = [&](uint32_t col, const float (&v)[16]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float a, b;
      ft_split<RN>(v[c], a, b);
      hi[c] = __float_as_uint(a), lo[c] = __float_as_uint(b);
    }
    tc::tmem_st16(tXhi + lane_off + col, hi);
    tc::tmem_st16(tXlo + lane_off + col, lo);
  };
  uint32_t ph = 0;
  // the tile's rows of h and agg -> shared memory, one tile AHEAD (under the previous tile's MMAs and epilogues)
  auto issue_loads = [&](int tile) {
    const int a0 = base + tile * FT_TILE;
    const int rows = tile < n_tiles ? min(FT_TILE, a_end - a0) : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = tid + FT_THREADS * k, r = i >> 3, c = i & 7;
      const bool ok = r < rows;
      const int64_t g = ok ? (int64_t)(a0 + r) * D + 4 * c : 0;
      const uint32_t dst = (uint32_t)((r * 8 + (c ^ (r & 7))) * 16), n = ok ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(s.in[0]) + dst), "l"(h + g), "r"(n) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(s.in[1]) + dst), "l"(agg + g), "r"(n) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto read_rows = [&](float (&hr)[16], float (&ar)[16]) {  // this thread's 16 columns of its row
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int o = (trow * 8 + ((4 * hf + c) ^ (trow & 7))) * 4;
      const float4 x = *reinterpret_cast<const float4*>(&s.in[0][o]), y = *reinterpret_cast<const float4*>(&s.in[1][o]);
      hr[4 * c] = x.x, hr[4 * c + 1] = x.y, hr[4 * c + 2] = x.z, hr[4 * c + 3] = x.w;
      ar[4 * c] = y.x, ar[4 * c + 1] = y.y, ar[4 * c + 2] = y.z, ar[4 * c + 3] = y.w;
    }
  };
  const int pair_id = 1 + q;  // named barrier of the two warps that share a quadrant (64 threads)
  const int t64 = hf * 32 + lane;
  // 16 columns of this thread's row -> the quadrant's staging rows -> global memory as whole 128-byte rows (the partner warp
  // holds the other 16 columns); rows_q = valid rows of this quadrant in the tile
  auto store_rows = [&](float* dst, int a0, int rows_q, const float (&v)[16]) {
    float* ob = s.out[q];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(&ob[(lane * 8 + ((4 * hf + c) ^ (lane & 7))) * 4]) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    tc::named_bar_sync(pair_id, 64);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = t64 + 64 * k, r = i >> 3, c = i & 7;
      const float4 x = *reinterpret_cast<const float4*>(&ob[(r * 8 + (c ^ (r & 7))) * 4]);
      if (r < rows_q) reinterpret_cast<float4*>(dst + (int64_t)(a0 + q * 32 + r) * D)[c] = x;
    }
    tc::named_bar_sync(pair_id, 64);  // the staging rows are free again
  };
  issue_loads(cta);
  for (int tile = cta; tile < n_tiles; tile += n_cta) {
    const int a0 = base + tile * FT_TILE;
    const int rows_q = min(FT_TILE, a_end - a0) - q * 32;  // valid rows of this quadrant (may be <= 0)
    float hv[16];
    {
      float an[16];
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();  // every thread's copies have landed
      read_rows(hv, an);
      to_tmem((uint32_t)cb, hv), to_tmem((uint32_t)(32 + cb), an);
    }
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_thread_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ko = (uint64_t)(ks * 2 * 64 * 16 / 16);
          ft_mma_ts(tD1, tXhi + 8 * ks, dW1[0] + ko, id64, ks > 0);
          ft_mma_ts(tD1, tXhi + 8 * ks, dW1[1] + ko, id64, true);
          ft_mma_ts(tD1, tXlo + 8 * ks, dW1[0] + ko, id64, true);
        }
        tc::mma_commit(&s.bar[0]);
      }
      __syncwarp();
    }
    issue_loads(tile + n_cta);  // next tile's rows (every thread has read this tile's: the barrier above)
    tc::mbar_wait(&s.bar[0], ph);
    tc::fence_after_thread_sync();
    float zv[16];
    {
      float v[16], rh[16];
      tc::tmem_ld16(tD1 + 32 + cb + lane_off, v);  // r
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        v[c] = ft_sigmoid(v[c] + s.br[cb + c]);
        rh[c] = v[c] * hv[c];
      }
      if (r_out) store_rows(r_out, a0, rows_q, v);
      to_tmem((uint32_t)cb, rh);  // over the h columns: the first product has consumed them
      tc::tmem_ld16(tD1 + cb + lane_off, v);  // z
#pragma unroll
      for (int c = 0; c < 16; ++c) zv[c] = ft_sigmoid(v[c] + s.bz[cb + c]);
    }
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_thread_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ko = (uint64_t)(ks * 2 * 32 * 16 / 16);
          ft_mma_ts(tD2, tXhi + 8 * ks, dW2[0] + ko, id32, ks > 0);
          ft_mma_ts(tD2, tXhi + 8 * ks, dW2[1] + ko, id32, true);
          ft_mma_ts(tD2, tXlo + 8 * ks, dW2[0] + ko, id32, true);
        }
        tc::mma_commit(&s.bar[1]);
      }
      __syncwarp();
    }
    if (z_out) store_rows(z_out, a0, rows_q, zv);
    tc::mbar_wait(&s.bar[1], ph);
    tc::fence_after_thread_sync();
    {
      float n[16], part = 0.f;
      tc::tmem_ld16(tD2 + cb + lane_off, n);
#pragma unroll
      for (int c = 0; c < 16; ++c) n[c] = ft_tanh(n[c] + s.bh[cb + c]);
      if (ht_out) store_rows(ht_out, a0, rows_q, n);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        n[c] = (1.0f - zv[c]) * hv[c] + zv[c] * n[c];
        part += n[c];
      }
      s.xs[0][hf][trow] = part;  // the other 16 columns of the row live in the partner warp
      tc::named_bar_sync(pair_id, 64);
      const float mean = (s.xs[0][0][trow] + s.xs[0][1][trow]) * (1.0f / D);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        n[c] -= mean;
        var = fmaf(n[c], n[c], var);
      }
      s.xs[1][hf][trow] = var;
      tc::named_bar_sync(pair_id, 64);
      const float inv = 1.0f / sqrtf((s.xs[1][0][trow] + s.xs[1][1][trow]) * (1.0f / D) + eps);
      {
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = n[c] * inv * s.gamma[cb + c] + s.beta[cb + c] + hv[c];
        store_rows(h_out, a0, rows_q, o);
      }
    }
    ph ^= 1;
    tc::fence_before_thread_sync();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tm);
}

}  // namespace imp

using namespace imp;

// d_z / d_r / d_ht: all three or none (NULL: plain forward)
extern "C" int imp_gated_update_tc32(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                     const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out,
                                     float* d_z, float* d_r, float* d_ht, void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG, "imp_gated_update_tc32: bad sizes");
  IMP_REQUIRE(d == FT_D, IMP_ERR_DIM, "imp_gated_update_tc32: atom_dim %d not supported (32)", d);
  if (n_atoms == 0) return 0;
  IMP_REQUIRE(d_h && d_agg && d_h_out && w_cat && w_an && w_cat->Wz && w_an->Wz, IMP_ERR_ARG, "imp_gated_update_tc32: null pointer");
  IMP_REQUIRE((d_z != nullptr) == (d_r != nullptr) && (d_z != nullptr) == (d_ht != nullptr), IMP_ERR_ARG,
              "imp_gated_update_tc32: pass all three gate outputs or none");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_gated_update_tc32: tcgen05 needs an sm_100 device");
  int dev = 0, sms = 148;
  IMP_CUDA(cudaGetDevice(&dev));
  IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles_cat = (int)ceil_div(n_cat_atoms, FT_TILE), tiles_an = (int)ceil_div(n_atoms - n_cat_atoms, FT_TILE);
  const int tiles = tiles_cat + tiles_an;
  const int grid = tiles < 2 * sms ? (tiles < 2 ? 2 : tiles) : 2 * sms;  // two CTAs per SM; both towers get at least one CTA
  int n_cta_cat = tiles > 0 ? (int)((int64_t)grid * tiles_cat / tiles) : 1;
  n_cta_cat = n_cta_cat < 1 ? 1 : (n_cta_cat > grid - 1 ? grid - 1 : n_cta_cat);
  const size_t smem = sizeof(FtSmem) + 1024;
  if (d_z) {  // training form: truncation splits
    IMP_CUDA(cudaFuncSetAttribute(gated_update_tc32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gated_update_tc32_kernel<false><<<grid, FT_THREADS, smem, (cudaStream_t)stream>>>(d_h, d_agg, n_atoms, n_cat_atoms, n_cta_cat, *w_cat,
                                                                                   *w_an, eps, d_h_out, d_z, d_r, d_ht);
  } else {  // inference form: rounded splits
    IMP_CUDA(cudaFuncSetAttribute(gated_update_tc32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gated_update_tc32_kernel<true><<<grid, FT_THREADS, smem, (cudaStream_t)stream>>>(d_h, d_agg, n_atoms, n_cat_atoms, n_cta_cat, *w_cat,
                                                                                  *w_an, eps, d_h_out, d_z, d_r, d_ht);
  }
  IMP_LAUNCH_CHECK();
  return 0;
}
