// GatedUpdate backward on the tensor cores (tcgen05, kind::tf32, fp32 accumulate in TMEM) for atom_dim 32.
//
// Replaces, for the training step (train_viscosity.py:227-230), what TensorFlow's autodiff does for GatedUpdate.call
// (models/layers.py:142-156): given dL/dh_out it returns dL/dh, dL/dagg and the gradients of the layer's eight variables.
// The forward kept the gates (imp_gated_update_train: z, r, tanh candidate), so nothing is recomputed; the six
// contractions that remain run as tcgen05.mma:
//     [dRH | dagg] = Gh . Wh^T                      (128 x 64, K = 32)        Gh = dL/d(candidate pre-activation)
//     [dh_zr | dagg_zr] = [Gz | Gr] . [Wz^T ; Wr^T]   (128 x 64, K = 64)        Gz, Gr = dL/d(gate pre-activations)
//     dW          += [h | agg | r*h | 1]^T . [Gz | Gr | Gh]   (97 x 96, K = 128 atoms; rows = dWz, dWr, dWh blocks and the biases)
// Gradients must stay within 2e-4 of fp64 autograd (tests/test_gpu_train.py), so every operand is split into two tf32
// terms x = hi + lo (hi = rna(x), lo = rna(x - hi)) and every product is three MMAs: hi.hi + hi.lo + lo.hi ("3xTF32":
// ~2^-21 relative per product).
//
// Mapping: one persistent CTA of 256 threads per SM; TWO threads per atom row of a 128-atom tile (TMEM lane = row; warps q and
// q + 4 share quadrant q and own 16 of the 32 columns each in every row-wise phase: eight warps in flight instead of four, half
// as long a dependent chain per warp; the row sums of the LayerNorm backward cross the pair through shared memory).  The row
// operands of the first three products (Gz, Gr, Gh; hi and lo) are written to TENSOR MEMORY by their owners
// (tcgen05.st, A operand of the "TS" MMA form); the weights are staged once per CTA as K-major hi / lo operands.  The
// weight-gradient product needs the ATOM index along K: each thread scatters its row into K-major [feature][atom]
// shared-memory operands (transposed staging, K-chunk stride padded by 16 bytes so that a warp's 32 stores hit 32 banks),
// half a tile (64 atoms) at a time; its accumulator (rows = features, 96 columns) stays in TMEM for ALL tiles of the CTA
// and is read once at the end.  The bias gradients come from a constant row of ones in that operand; dgamma / dbeta are
// column sums taken with a 31-shuffle transpose-reduce per warp.  Every sum has a fixed order: bit-reproducible.
#include "common.cuh"
#include "tc_common.cuh"

namespace imp {

constexpr int BT_D = 32;
constexpr int BT_TILE = 128;
constexpr int BT_HALF = 64;                         // atoms staged per weight-gradient pass
constexpr int BT_AROWS = 96;                        // feature rows written (h, agg, r*h); row 96 = ones; M = 128 reads on
constexpr int BT_LBO = (BT_AROWS + 1) * 16;         // bytes between K chunks (4 atoms): 97 rows of 16 B
constexpr int BT_OPBYTES = (BT_HALF / 4) * BT_LBO;  // one staged operand (hi or lo) of a half tile
static_assert(BT_LBO % 16 == 0 && ((BT_LBO / 4) % 32) == 4, "chunk stride: 16-byte aligned, 4 banks apart: a warp's 32 stores hit 32 banks");

struct BtSmem {
  // K-major tf32 operands of the weights, hi then lo: W1 [64 x 32] = Wh; W2 [64 x 64]: row n = input n of [h | agg], k < 32: Wz, else Wr
  float W1[2][64 * 32];
  float W2[2][64 * 64];
  unsigned char stage[4 * BT_OPBYTES + 2048];  // A_hi, A_lo, B_hi, B_lo of a half tile (+ slack: M = 128 reads past row 96)
  float gamma[BT_D];
  float red[8][BT_D];            // per warp: dgamma (lanes 0-15) and dbeta (16-31) column sums of its 16 columns
  float xs[4][2][BT_TILE];       // row sums exchanged between the two column halves of a row
  // the next tile's h, g_out, z and tanh-candidate rows, copied by cp.async with 8 lanes per 128-byte row (a thread that loads
  // its own row touches 32 L1 lines per instruction: with eight arrays per tile that, not HBM, was half of a tile's time);
  // 16-byte chunk c of row r at chunk c ^ (r & 7): the row owners read their chunks back without bank conflicts.  in[0] doubles
  // as the staging rows of the two outputs (dh, dagg), which leave as whole rows too.
  float in[4][BT_TILE * BT_D];
  uint64_t bar[4];  // 0: B1 done, 1: B2 done, 2: weight-gradient MMAs of the staged half done
  uint32_t tmem_base;
};

__device__ __forceinline__ float bt_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// x = hi + lo with hi a tf32 number: hi = x with its 13 low significand bits cleared (one LOP3), lo = x - hi (exact in fp32).
// The MMA ignores the 13 low bits of lo, i.e. lo is truncated to tf32 by the hardware: |x - hi - lo_tf32| <= 2^-21 |x|.
// (cvt.rna on both terms gave the same accuracy at three instructions per value, two of them conversions.)
__device__ __forceinline__ void bt_split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void bt_mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}
// lane j of the warp receives sum over the warp's 32 lanes of v[j] (fixed order: bit-reproducible)
__device__ __forceinline__ float bt_column_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (upper ? v[i + s] : v[i]) + recv;
    }
  }
  return v[0];
}

// 16 columns over the warp's 32 rows: lane j and lane j + 16 receive the sum over the lanes of v[j & 15] (fixed order)
__device__ __forceinline__ float bt_column_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool upper = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (upper ? v[i + s] : v[i]) + recv;
    }
  }
  return v[0];
}

constexpr int BT_THREADS = 2 * BT_TILE;  // two threads per atom row: warps q and q + 4 share TMEM quadrant q, 16 columns each

__global__ void __launch_bounds__(BT_THREADS, 1) gated_update_bwd_tc_kernel(
    const float* __restrict__ h, const float* __restrict__ agg, const float* __restrict__ zs, const float* __restrict__ rs,
    const float* __restrict__ hts, const float* __restrict__ g_out, int n_atoms, int n_cat, int n_cta_cat, imp_gru_weights_t wc,
    imp_gru_weights_t wa, float eps, float* __restrict__ dh, float* __restrict__ dagg, float* __restrict__ partial) {
  constexpr int D = BT_D;
  extern __shared__ __align__(1024) unsigned char bt_raw[];
  BtSmem& s = *reinterpret_cast<BtSmem*>(bt_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;  // TMEM quadrant; column half of this thread
  const int trow = q * 32 + lane;          // row of the tile = TMEM lane
  const int cb = (BT_D / 2) * hf;          // first of the 16 columns this thread owns in every row-wise phase
  const bool is_cat = (int)blockIdx.x < n_cta_cat;
  const imp_gru_weights_t& w = is_cat ? wc : wa;
  const int base = is_cat ? 0 : n_cat, a_end = is_cat ? n_cat : n_atoms;
  const int n_tiles = (a_end - base + BT_TILE - 1) / BT_TILE;
  const int cta = is_cat ? blockIdx.x : blockIdx.x - n_cta_cat, n_cta = is_cat ? n_cta_cat : gridDim.x - n_cta_cat;

  // ---- weights -> K-major tf32 hi / lo operands (element (n, k) at chunk_off(n, k / 4, R) + (k % 4) * 4)
  for (int i = tid; i < 64 * 32; i += BT_THREADS) {
    {  // W1[n][k] = Wh[n][k]   (dX[a][n] = sum_j Gh[a][j] Wh[n][j]; Wh is [2d in][d out] row-major)
      const int n = i / 32, k = i % 32;
      float hi, lo;
      bt_split(__ldg(w.Wh + n * D + k), hi, lo);
      const int o = (tc::chunk_off(n, k / 4, 64) + (k % 4) * 4) / 4;
      s.W1[0][o] = hi, s.W1[1][o] = lo;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {  // W2[n][k] = (k < 32 ? Wz : Wr)[n][k & 31], n = input index of [h | agg]
      const int n = i / 64 + 32 * half, k = i % 64;
      const float* src = k < 32 ? w.Wz : w.Wr;
      float hi, lo;
      bt_split(__ldg(src + n * D + (k & 31)), hi, lo);
      const int o = (tc::chunk_off(n, k / 4, 64) + (k % 4) * 4) / 4;
      s.W2[0][o] = hi, s.W2[1][o] = lo;
    }
  }
  if (tid < D) s.gamma[tid] = w.gamma[tid];
  // the constant row of ones of the transposed operand (row 96; its lo term is zero) and zeros elsewhere in the slack
  for (int i = tid; i < (int)sizeof(s.stage) / 4; i += BT_THREADS) reinterpret_cast<float*>(s.stage)[i] = 0.f;
  __syncthreads();
  for (int k = tid; k < BT_HALF; k += BT_THREADS)
    *reinterpret_cast<float*>(s.stage + (k / 4) * BT_LBO + BT_AROWS * 16 + (k % 4) * 4) = 1.0f;
  if (warp == 0) tc::tmem_alloc<512>(&s.tmem_base);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::mbar_init(&s.bar[i], 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();

  const uint32_t tm = s.tmem_base, lane_off = (uint32_t)(q * 32) << 16;
  // G: [Gz | Gr | Gh] hi, lo; D: [dRH | dagg_h] of B1; D2: [dh_zr | dagg_zr] of B2; DW: the weight-gradient accumulator
  const uint32_t tGhi = tm, tGlo = tm + 96, tD = tm + 192, tD2 = tm + 256, tDW = tm + 320;
  const uint32_t id64 = tc::make_idesc(tc::FMT_TF32, BT_TILE, 64);
  const uint32_t id96 = tc::make_idesc(tc::FMT_TF32, BT_TILE, 96);
  const uint64_t dW1[2] = {tc::make_smem_desc(tc::smem_u32(s.W1[0]), 64 * 16, 128), tc::make_smem_desc(tc::smem_u32(s.W1[1]), 64 * 16, 128)};
  const uint64_t dW2[2] = {tc::make_smem_desc(tc::smem_u32(s.W2[0]), 64 * 16, 128), tc::make_smem_desc(tc::smem_u32(s.W2[1]), 64 * 16, 128)};
  unsigned char* sAh = s.stage;
  unsigned char* sAl = s.stage + BT_OPBYTES;
  unsigned char* sBh = s.stage + 2 * BT_OPBYTES;
  unsigned char* sBl = s.stage + 3 * BT_OPBYTES;
  const uint64_t dA[2] = {tc::make_smem_desc(tc::smem_u32(sAh), BT_LBO, 128), tc::make_smem_desc(tc::smem_u32(sAl), BT_LBO, 128)};
  const uint64_t dB[2] = {tc::make_smem_desc(tc::smem_u32(sBh), BT_LBO, 128), tc::make_smem_desc(tc::smem_u32(sBl), BT_LBO, 128)};

  float agam = 0.f, abet = 0.f;  // lane j: column cb + (j & 15) of dgamma / dbeta over this warp's rows, all tiles
  uint32_t ph01 = 0, ph2 = 0;    // mbarrier parities (bars 0 and 1 flip once per tile, bar 2 twice)
  bool dw_started = false, dw_pending = false;
  const int pair_id = 1 + q;     // named barrier of the two warps that share a row (64 threads)

  // row of the half tile -> byte offset of element (feature row f, atom k = trow % 64) inside a staged operand
  const int kk = trow & (BT_HALF - 1);
  const uint32_t st_off = (uint32_t)((kk / 4) * BT_LBO + (kk % 4) * 4);
  auto stage_vec = [&](unsigned char* hi_base, unsigned char* lo_base, int row0, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float hi, lo;
      bt_split(v[c], hi, lo);
      *reinterpret_cast<float*>(hi_base + st_off + (row0 + c) * 16) = hi;
      *reinterpret_cast<float*>(lo_base + st_off + (row0 + c) * 16) = lo;
    }
  };
  auto stage_raw = [&](unsigned char* dst, int row0, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) *reinterpret_cast<float*>(dst + st_off + (row0 + c) * 16) = v[c];
  };
  auto to_tmem = [&](uint32_t col, const float (&v)[16]) {  // hi / lo halves of a 16-column block of the row operand
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float a, b;
      bt_split(v[c], a, b);
      hi[c] = __float_as_uint(a), lo[c] = __float_as_uint(b);
    }
    tc::tmem_st16(tGhi + lane_off + col, hi);
    tc::tmem_st16(tGlo + lane_off + col, lo);
  };
  auto load_row = [&](const float* p, int row, bool ok, float (&v)[16]) {  // this thread's 16 columns
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 x = ok ? __ldg(reinterpret_cast<const float4*>(p + (int64_t)row * D + cb) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * c] = x.x, v[4 * c + 1] = x.y, v[4 * c + 2] = x.z, v[4 * c + 3] = x.w;
    }
  };
  // sum of a per-thread partial over the two column halves of the row (the partner warp holds the other half)
  auto row_sum = [&](int slot, float part) {
    s.xs[slot][hf][trow] = part;
    tc::named_bar_sync(pair_id, 64);
    return s.xs[slot][0][trow] + s.xs[slot][1][trow];
  };

  // arrays [first, last) of {h, g_out, z, tanh candidate} of a tile -> shared memory
  auto issue_in = [&](int tile, int first, int last) {
    const int a0 = base + tile * BT_TILE;
    const int rows = tile < n_tiles ? min(BT_TILE, a_end - a0) : 0;
    const float* src[4] = {h, g_out, zs, hts};
#pragma unroll
    for (int arr = 0; arr < 4; ++arr) {
      if (arr < first || arr >= last) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = tid + BT_THREADS * k, r = i >> 3, c = i & 7;
        const bool okr = r < rows;
        const int64_t g = okr ? (int64_t)(a0 + r) * D + 4 * c : 0;
        const uint32_t dst = tc::smem_u32(s.in[arr]) + (uint32_t)((r * 8 + (c ^ (r & 7))) * 16), n = okr ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src[arr] + g), "r"(n) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto read_in = [&](int arr, float (&v)[16]) {  // this thread's 16 columns of its row
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 x = *reinterpret_cast<const float4*>(&s.in[arr][(trow * 8 + ((4 * hf + c) ^ (trow & 7))) * 4]);
      v[4 * c] = x.x, v[4 * c + 1] = x.y, v[4 * c + 2] = x.z, v[4 * c + 3] = x.w;
    }
  };
  const int t64 = hf * 32 + lane;
  // 16 columns of this thread's row -> the quadrant's staging rows (inside in[0]) -> global memory as whole 128-byte rows
  auto store_rows = [&](float* dst, int a0, int rows_q, const float (&v)[16]) {
    float* ob = s.in[0] + q * 32 * D;
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
  issue_in(cta, 0, 4);
  for (int tile = cta; tile < n_tiles; tile += n_cta) {
    const int a0 = base + tile * BT_TILE;
    const int row = a0 + trow;
    const bool ok = trow < min(BT_TILE, a_end - a0);
    const int rows_q = min(BT_TILE, a_end - a0) - q * 32;  // valid rows of this quadrant (may be <= 0)
    {  // the next tile's r and agg rows (read directly) -> L2
      const int nrow = row + n_cta * BT_TILE;
      if (nrow < a_end) asm volatile("prefetch.global.L2 [%0];" ::"l"((hf ? agg : rs) + (int64_t)nrow * D));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // every thread's copies of this tile's h, g_out, z, tanh-candidate rows have landed
    float hv[16], go[16];
    read_in(0, hv), read_in(1, go);
    {  // LayerNorm forward statistics and backward, gate gradients (models/layers.py:151-156 under autodiff)
      float zv[16], tv[16], gz[16], gh[16], gx[16];
      read_in(2, zv), read_in(3, tv);
      float nrm[16], part = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        nrm[c] = fmaf(zv[c], tv[c] - hv[c], hv[c]);
        part += nrm[c];
      }
      const float mean = row_sum(0, part) * (1.0f / D);
      part = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        nrm[c] -= mean;
        part = fmaf(nrm[c], nrm[c], part);
      }
      const float inv = 1.0f / sqrtf(row_sum(1, part) * (1.0f / D) + eps);
      float p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        nrm[c] *= inv;  // xhat
        gx[c] = go[c] * nrm[c];
        const float dx = go[c] * s.gamma[cb + c];
        p1 += dx;
        p2 = fmaf(dx, nrm[c], p2);
      }
      s.xs[3][hf][trow] = p2;
      const float m1 = row_sum(2, p1) * (1.0f / D);  // (the barrier inside also publishes slot 3)
      const float m2 = (s.xs[3][0][trow] + s.xs[3][1][trow]) * (1.0f / D);
      float gbeta[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        gbeta[c] = go[c];
        const float dn = inv * (go[c] * s.gamma[cb + c] - m1 - nrm[c] * m2);
        const float z = zv[c], ht = tv[c], hj = hv[c];
        go[c] = fmaf(dn, 1.0f - z, go[c]);        // dh: residual path + the (1 - z) path
        gz[c] = dn * (ht - hj) * z * (1.0f - z);  // dL/dzpre
        gh[c] = dn * z * (1.0f - ht * ht);        // dL/dhpre
      }
      agam += bt_column_sum16(gx, lane);
      abet += bt_column_sum16(gbeta, lane);
      to_tmem((uint32_t)cb, gz), to_tmem((uint32_t)(64 + cb), gh);
    }
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) {  // B1: [dRH | dagg] = Gh . Wh^T, K = 32
      tc::fence_after_thread_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ko = (uint64_t)(ks * 2 * 64 * 16 / 16);
          bt_mma_ts(tD, tGhi + 64 + 8 * ks, dW1[0] + ko, id64, ks > 0);
          bt_mma_ts(tD, tGhi + 64 + 8 * ks, dW1[1] + ko, id64, true);
          bt_mma_ts(tD, tGlo + 64 + 8 * ks, dW1[0] + ko, id64, true);
        }
        tc::mma_commit(&s.bar[0]);
      }
      __syncwarp();
    }
    issue_in(tile + n_cta, 1, 4);  // next tile's g_out, z, tanh candidate (every thread has read this tile's: the barrier above)
    float rv[16];
    load_row(rs, row, ok, rv);
    // while B1 runs: the staging buffers are free once the previous tile's second weight-gradient pass has been consumed;
    // the warps of the first half tile scatter the operand rows that do not depend on B1 (h, agg)
    if (dw_pending) {
      tc::mbar_wait(&s.bar[2], ph2);
      ph2 ^= 1;
      dw_pending = false;
    }
    if ((q >> 1) == 0) {
      float v[16];
      stage_vec(sAh, sAl, cb, hv);
      load_row(agg, row, ok, v);
      stage_vec(sAh, sAl, 32 + cb, v);
    }
    tc::mbar_wait(&s.bar[0], ph01);
    tc::fence_after_thread_sync();
    float rh[16];
    {
      float drh[16], gr[16];
      tc::tmem_ld16(tD + lane_off + cb, drh);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        go[c] = fmaf(drh[c], rv[c], go[c]);
        gr[c] = drh[c] * hv[c] * rv[c] * (1.0f - rv[c]);  // dL/drpre
        rh[c] = rv[c] * hv[c];
      }
      to_tmem((uint32_t)(32 + cb), gr);
    }
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if ((q >> 1) == half) {  // the warps that own these 64 atoms scatter their rows: [feature][atom], hi and lo
        float v[16];
        if (half == 1) {  // (the first half's h and agg rows were staged while B1 ran)
          stage_vec(sAh, sAl, cb, hv);
          load_row(agg, row, ok, v);
          stage_vec(sAh, sAl, 32 + cb, v);
        }
        stage_vec(sAh, sAl, 64 + cb, rh);
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {  // Gz, Gr, Gh: their hi / lo terms are in tensor memory already
          tc::tmem_ld16(tGhi + lane_off + 32 * blk + cb, v);
          stage_raw(sBh, 32 * blk + cb, v);
          tc::tmem_ld16(tGlo + lane_off + 32 * blk + cb, v);
          stage_raw(sBl, 32 * blk + cb, v);
        }
      }
      tc::fence_proxy_async_smem();
      __syncthreads();
      if (half == 1) issue_in(tile + n_cta, 0, 1);  // next tile's h: its buffer staged this tile's outputs until the barrier above
      if (warp == 0) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
          if (half == 0) {  // B2: [dh_zr | dagg_zr] = [Gz | Gr] . W2^T, K = 64 (its own columns: dagg_h of B1 is added in registers)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t ko = (uint64_t)(ks * 2 * 64 * 16 / 16);
              bt_mma_ts(tD2, tGhi + 8 * ks, dW2[0] + ko, id64, ks > 0);
              bt_mma_ts(tD2, tGhi + 8 * ks, dW2[1] + ko, id64, true);
              bt_mma_ts(tD2, tGlo + 8 * ks, dW2[0] + ko, id64, true);
            }
            tc::mma_commit(&s.bar[1]);
          }
#pragma unroll
          for (int ks = 0; ks < BT_HALF / 8; ++ks) {  // dW += A^T-operand . B^T-operand over these 64 atoms
            const uint64_t ko = (uint64_t)(ks * 2 * BT_LBO / 16);
            tc::mma_tf32(tDW, dA[0] + ko, dB[0] + ko, id96, dw_started || ks > 0);
            tc::mma_tf32(tDW, dA[0] + ko, dB[1] + ko, id96, true);
            tc::mma_tf32(tDW, dA[1] + ko, dB[0] + ko, id96, true);
          }
          tc::mma_commit(&s.bar[2]);
        }
        __syncwarp();
      }
      dw_started = true;
      if (half == 0) {
        tc::mbar_wait(&s.bar[1], ph01);
        tc::fence_after_thread_sync();
        float v[16], u[16];
        tc::tmem_ld16(tD2 + lane_off + cb, v);  // dh_zr
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] += go[c];
        store_rows(dh, a0, rows_q, v);
        tc::tmem_ld16(tD + 32 + lane_off + cb, u);   // dagg through Wh
        tc::tmem_ld16(tD2 + 32 + lane_off + cb, v);  // dagg through Wz, Wr
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] += u[c];
        store_rows(dagg, a0, rows_q, v);
        tc::mbar_wait(&s.bar[2], ph2);  // the first half has been consumed: the other warps may overwrite the buffers
        ph2 ^= 1;
      } else {
        dw_pending = true;
      }
    }
    ph01 ^= 1;
    tc::fence_before_thread_sync();
  }
  if (dw_pending) tc::mbar_wait(&s.bar[2], ph2);
  tc::fence_after_thread_sync();

  // ---- per-CTA partial gradients, layout [dWz (2d, d) | dbz | dWr | dbr | dWh | dbh | dgamma | dbeta]
  float* o = partial + (int64_t)blockIdx.x * (3 * 2 * D * D + 5 * D);
  constexpr int BLK = 2 * D * D + D;
  if (dw_started) {
    float v[16];
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {  // columns [32 blk + cb, + 16): x^T Gz, x^T Gr, x^T Gh  (all lanes load: .sync.aligned)
      tc::tmem_ld16(tDW + lane_off + 32 * blk + cb, v);
      if (trow == BT_AROWS) {  // the ones row: bias gradients
#pragma unroll
        for (int j = 0; j < 16; ++j) o[blk * BLK + 2 * D * D + cb + j] = v[j];
      } else if (blk < 2) {
        if (trow < 2 * D) {  // rows h (0..31) and agg (32..63): dWz / dWr rows k = trow
#pragma unroll
          for (int j = 0; j < 16; ++j) o[blk * BLK + trow * D + cb + j] = v[j];
        }
      } else if (trow >= D && trow < BT_AROWS) {  // dWh: rows k < d multiply r*h (feature rows 64..95), rows d + k multiply agg (32..63)
        const int k = trow >= 2 * D ? trow - 2 * D : trow;
#pragma unroll
        for (int j = 0; j < 16; ++j) o[2 * BLK + k * D + cb + j] = v[j];
      }
    }
  } else {
    for (int i = tid; i < 3 * BLK; i += BT_THREADS) o[i] = 0.f;
  }
  if (lane < 16) s.red[warp][lane] = agam, s.red[warp][16 + lane] = abet;
  __syncthreads();
  if (tid < 2 * D) {  // tid < 32: dgamma column tid, else dbeta column tid - 32; the column's half lives in warps 4 (col / 16) + q
    const int col = tid & (D - 1), w0 = 4 * (col >> 4), j = (col & 15) + (tid >= D ? 16 : 0);
    o[3 * BLK + tid] = (s.red[w0][j] + s.red[w0 + 1][j]) + (s.red[w0 + 2][j] + s.red[w0 + 3][j]);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tm);
}

__global__ void bt_reduce_partials_kernel(const float* __restrict__ partial, int n_parts, int64_t stride, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int p = 0; p < n_parts; ++p) acc += partial[(int64_t)p * stride + i];
  out[i] = acc;
}

}  // namespace imp

using namespace imp;

extern "C" int imp_gated_update_bwd_tc(const float* d_h, const float* d_agg, const float* d_z, const float* d_r, const float* d_ht,
                                       const float* d_gout, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                       const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_dh,
                                       float* d_dagg, float* d_grads_cat, float* d_grads_an, float* d_workspace, void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG, "imp_gated_update_bwd_tc: bad sizes");
  IMP_REQUIRE(d == BT_D, IMP_ERR_DIM, "imp_gated_update_bwd_tc: atom_dim %d not supported (32)", d);
  IMP_REQUIRE(d_h && d_agg && d_z && d_r && d_ht && d_gout && d_dh && d_dagg && d_grads_cat && d_grads_an && d_workspace && w_cat && w_an,
              IMP_ERR_ARG, "imp_gated_update_bwd_tc: null pointer");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_gated_update_bwd_tc: tcgen05 needs an sm_100 device");
  int dev = 0, sms = 148;
  IMP_CUDA(cudaGetDevice(&dev));
  IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles_cat = (int)ceil_div(n_cat_atoms, BT_TILE), tiles_an = (int)ceil_div(n_atoms - n_cat_atoms, BT_TILE);
  int n_cat = tiles_cat + tiles_an > 0 ? (int)((int64_t)sms * tiles_cat / (tiles_cat + tiles_an)) : 1;
  n_cat = n_cat < 1 ? 1 : (n_cat > sms - 1 ? sms - 1 : n_cat);
  const int grid = sms;
  const size_t smem = sizeof(BtSmem) + 1024;
  IMP_CUDA(cudaFuncSetAttribute(gated_update_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  gated_update_bwd_tc_kernel<<<grid, BT_THREADS, smem, st>>>(d_h, d_agg, d_z, d_r, d_ht, d_gout, n_atoms, n_cat_atoms, n_cat, *w_cat, *w_an,
                                                         eps, d_dh, d_dagg, d_workspace);
  IMP_LAUNCH_CHECK();
  const int n = 3 * 2 * BT_D * BT_D + 5 * BT_D;
  bt_reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_workspace, n_cat, n, n, d_grads_cat);
  IMP_LAUNCH_CHECK();
  bt_reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_workspace + (int64_t)n_cat * n, grid - n_cat, n, n, d_grads_an);
  IMP_LAUNCH_CHECK();
  return 0;
}
