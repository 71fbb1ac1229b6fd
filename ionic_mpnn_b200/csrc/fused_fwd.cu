// Whole-tower fused MPNN forward on sm_100a: Embedding -> [BondMatrixMessage o Reduce -> GatedUpdate] x S ->
// GlobalSumPool in ONE persistent kernel, atom states resident on chip for all S steps.
//
// Replaces, for atom_dim 32 / bond_dim 8 / <= 4 steps (the viscosity model of train_viscosity.py:139-231):
//   Embedding(atom)                     train_viscosity.py:163,171
//   BondMatrixMessage.call + Reduce     models/layers.py:100-117, 57-83
//   GatedUpdate.call                    models/layers.py:142-156
//   GlobalSumPool.call                  models/layers.py:161-164
//
// Why it can be fused: no edge crosses a molecule, so a tile of whole molecules (<= 128 atoms) is closed under
// message passing.  HBM traffic drops from ~5 activation round trips per step to the index stream only
// (~23 B/atom + 8 B/entry); the kernel is bound by the SM's FP32 issue rate, not by memory.
//
// Re-association that turns the per-edge mat-vec into a dense GEMM.  With A_e = sum_k c_e[k] W_k
// (models/layers.py:108, c_e = bond embedding row) the aggregated message of destination v is
//     agg[v][l] = sum_e mult_e sum_m A_e[l][m] h[src_e][m]
//               = sum_{m,k} W[k][l][m] * Z[v][m*8+k],      Z[v][m*8+k] = sum_e mult_e c_e[k] h[src_e][m]
// so agg = Z (128 x 256) . Wc (256 x 32).  Z is built by the thread that owns row v (fp32 FMA, CSR order,
// deterministic), rounded to 16 bits and written straight into TENSOR MEMORY (tcgen05.st); all GEMMs take their A
// operand from TMEM (tcgen05.mma "TS" form) and their B operand (pre-packed weights, resident for all S steps) from
// shared memory, accumulate in fp32 in TMEM, and are read back with tcgen05.ld so that one thread owns one atom row:
// gates, blend, LayerNorm and residual need no cross-thread traffic.
//
// CTA = 256 threads = 2 independent warpgroups, each running its own 128-row tile pipeline on its own 256 TMEM
// columns, sharing the tower's weights in shared memory.  One CTA per SM, persistent; CTAs [0, n_cta_cat) serve the
// cation tower, the rest the anion tower (weights are per tower: train_viscosity.py:176-189).
//
// TMEM columns of a warpgroup (base = 256 * wg):
//   [  0,128)  Z, 16-bit pairs (K = 256)        -- dead after GEMM1, then reused:
//   [  0, 16)  h   operand (K = 32)   [ 16, 32)  agg operand    [ 32, 48)  r*h operand
//   [128,160)  GEMM1 accumulator (agg, fp32)    [160,224)  z | r pre-activations    [224,256)  candidate pre-activation
#include "fused_common.cuh"

namespace imp {

// KHALF = false: Wc[n = l][kk = m*8 + k] = W[k][l][m] as one [32 x 256] K-major block (first-generation kernels).
// KHALF = true : two [32 x 128] blocks, block hz holds k in [4 hz, 4 hz + 4): Wc_hz[l][m*4 + (k - 4 hz)] = W[k][l][m]
//                (the four-context kernel builds and multiplies Z in two K halves).
// wh_top_scale: generation 4 packs Wh[0:d] (the rows that multiply r * h) pre-multiplied by 0.5 (fused_fwd4.cu).
template <int FMT, bool KHALF>
__global__ void fused_pack_kernel(const float* __restrict__ W /* [K, d, d] */, imp_gru_weights_t w,
                                  unsigned char* __restrict__ out, float wh_top_scale) {
  constexpr int D = FZ_D, KK = FZ_D * FZ_K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D * KK) {
    const int l = i / KK, kk = i % KK, m = kk / FZ_K, k = kk % FZ_K;
    if (KHALF) {
      const int hz = k / 4, kq = m * 4 + (k % 4);
      *reinterpret_cast<uint16_t*>(out + hz * (FusedPack::WC_BYTES / 2) + tc::chunk_off(l, kq / 8, D) + (kq % 8) * 2) =
          tc::cvt16<FMT>(W[(k * D + l) * D + m]);
    } else {
      *reinterpret_cast<uint16_t*>(out + tc::chunk_off(l, kk / 8, D) + (kk % 8) * 2) = tc::cvt16<FMT>(W[(k * D + l) * D + m]);
    }
  }
  // third generation: sigmoid(x) = 0.5 tanh(0.5 x) + 0.5 with the 0.5 folded into [Wz | Wr] and their biases
  const float gate_scale = KHALF ? 0.5f : 1.0f;
  if (i < 2 * D * 2 * D) {  // Bzr[n][k] = (n < D ? Wz[k][n] : Wr[k][n - D])
    const int n = i / (2 * D), k = i % (2 * D);
    const float v = gate_scale * (n < D ? w.Wz[k * D + n] : w.Wr[k * D + (n - D)]);
    *reinterpret_cast<uint16_t*>(out + FusedPack::OFF_BZR + tc::chunk_off(n, k / 8, 2 * D) + (k % 8) * 2) = tc::cvt16<FMT>(v);
  }
  if (i < 2 * D * 16) {  // bias block of the gates: element (n, k) of a [64 x 16] K-major tile, only k = 0 is non-zero
    const int n = i / 16, k = i % 16;
    const float v = k == 0 ? gate_scale * (n < D ? w.bz[n] : w.br[n - D]) : 0.f;
    *reinterpret_cast<uint16_t*>(out + FusedPack::OFF_BBZR + tc::chunk_off(n, k / 8, 2 * D) + (k % 8) * 2) = tc::cvt16<FMT>(v);
  }
  if (i < D * 16) {  // bias block of the candidate
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<uint16_t*>(out + FusedPack::OFF_BBH + tc::chunk_off(n, k / 8, D) + (k % 8) * 2) =
        tc::cvt16<FMT>(k == 0 ? w.bh[n] : 0.f);
  }
  if (i < D * 2 * D) {  // Bh[n][k] = Wh[k][n]
    const int n = i / (2 * D), k = i % (2 * D);
    *reinterpret_cast<uint16_t*>(out + FusedPack::OFF_BH + tc::chunk_off(n, k / 8, D) + (k % 8) * 2) =
        tc::cvt16<FMT>((k < D ? wh_top_scale : 1.0f) * w.Wh[k * D + n]);
  }
  if (i < D) {
    float* b = reinterpret_cast<float*>(out + FusedPack::OFF_BIAS);
    b[i] = w.bz[i], b[D + i] = w.br[i], b[2 * D + i] = w.bh[i], b[3 * D + i] = w.gamma[i], b[4 * D + i] = w.beta[i];
  }
}


struct alignas(16) FusedWgSmem {
  float h[FZ_ROWS * FZ_HS];
  int molp[FZ_GROUP + 4];
  unsigned char amask[FZ_ROWS];
  uint64_t bar[4];
};


__host__ __device__ inline int fused_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * FZ_CS * 4 + 127) / 128 * 128;
  return steps * FusedPack::BYTES + ctab + 2 * (int)sizeof(FusedWgSmem) + (int)sizeof(FusedCtl);
}

// MP = number of h columns (m) handled per pass over a row's entries: MP*8 fp32 accumulators live in registers.
template <int FMT, bool PRECISE, int MP>
__global__ void __launch_bounds__(256, 1) mpnn_fused_kernel(const FusedArgs a) {
  static_assert(MP == 8 || MP == 16, "MP");
  constexpr int D = FZ_D;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = tid >> 7, t = tid & 127;
  const int wbytes = a.steps * FusedPack::BYTES;
  const int ctab_bytes = (a.bond_vocab * FZ_CS * 4 + 127) / 128 * 128;
  float* s_ctab = reinterpret_cast<float*>(smem + wbytes);
  FusedWgSmem& ws = reinterpret_cast<FusedWgSmem*>(smem + wbytes + ctab_bytes)[wg];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + 2 * sizeof(FusedWgSmem));

  const int tower = blockIdx.x >= a.n_cta_cat;
  {  // resident weights of this tower (all steps), bond coefficients
    const uint4* src = reinterpret_cast<const uint4*>(a.packed + (size_t)tower * wbytes);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < wbytes / 16; i += 256) dst[i] = __ldg(src + i);
    for (int i = tid; i < a.bond_vocab * FZ_K; i += 256) s_ctab[(i / FZ_K) * FZ_CS + (i % FZ_K)] = __ldg(a.bond_emb + i);
  }
  if (warp == 0) tc::tmem_alloc<512>(&ctl.tmem_base);
  if (t == 0) {
    tc::mbar_init(&ws.bar[0], 1);
    tc::mbar_init(&ws.bar[1], 1);
    tc::mbar_init(&ws.bar[2], 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();

  const uint32_t tbase = ctl.tmem_base + (uint32_t)(wg * 256);
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tZ = tbase, tAh = tbase, tAagg = tbase + 16, tArh = tbase + 32;
  const uint32_t tCagg = tbase + 128, tCzr = tbase + 160, tCht = tbase + 224;
  const uint32_t idesc32 = tc::make_idesc(FMT, FZ_ROWS, D), idesc64 = tc::make_idesc(FMT, FZ_ROWS, 2 * D);
  const uint32_t sw0 = tc::smem_u32(smem);
  const int bar_id = 1 + wg;

  const int P = a.n_pairs;
  const int n_groups = (P + FZ_GROUP - 1) / FZ_GROUP;
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  float* hrow = &ws.h[t * FZ_HS];
  uint32_t ph = 0;

  for (int g = cta_in_tower * 2 + wg; g < n_groups; g += n_cta_tower * 2) {
    const int m0 = g * FZ_GROUP, nm = min(FZ_GROUP, P - m0);
    const int base_mol = tower * P + m0;
    tc::named_bar_sync(bar_id, 128);
    if (t <= nm) ws.molp[t] = __ldg(a.mol_ptr + base_mol + t);
    tc::named_bar_sync(bar_id, 128);
    int ms = 0;
    while (ms < nm) {
      const int a0 = ws.molp[ms];
      int me = ms + 1;
      while (me < nm && ws.molp[me + 1] - a0 <= FZ_ROWS) ++me;
      int rows = ws.molp[me] - a0;
      if (rows > FZ_ROWS) {  // a single molecule larger than a tile: flagged, never read out of bounds
        if (t == 0 && a.status) *a.status = 1;
        rows = FZ_ROWS;
      }
      // ---------------------------------------------------------------- Embedding(atom)
      const bool valid = t < rows;
      int aid = 0, e0 = 0, e1 = 0;
      if (valid) {
        aid = __ldg(a.atom_id + a0 + t);
        e0 = __ldg(a.row_ptr + a0 + t);
        e1 = __ldg(a.row_ptr + a0 + t + 1);
      }
      {
        const int id = min(max(aid, 0), a.atom_vocab - 1);
#pragma unroll
        for (int c = 0; c < D / 4; ++c)
          reinterpret_cast<float4*>(hrow)[c] = valid ? __ldg(emb4 + id * (D / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        ws.amask[t] = (valid && aid > 0) ? 1 : 0;  // models/layers.py:163
      }
      tc::named_bar_sync(bar_id, 128);

      for (int s = 0; s < a.steps; ++s) {
        const uint32_t sw = sw0 + (uint32_t)(s * FusedPack::BYTES);
        const float* bias = reinterpret_cast<const float*>(smem + s * FusedPack::BYTES + FusedPack::OFF_BIAS);
        // ------------------------------------------------------------ Z rows -> TMEM
#pragma unroll 1
        for (int pass = 0; pass < D / MP; ++pass) {
          float acc[MP * FZ_K];
#pragma unroll
          for (int i = 0; i < MP * FZ_K; ++i) acc[i] = 0.f;
#pragma unroll 1
          for (int e = e0; e < e1; ++e) {
            const int bm = __ldg(a.edge_bm + e);
            int src = __ldg(a.col_src + e) - a0;
            src = min(max(src, 0), FZ_ROWS - 1);
            const float mult = (float)(bm >> 16);
            const int bond = min(bm & 0xffff, a.bond_vocab - 1);
            const float4 ca = *reinterpret_cast<const float4*>(s_ctab + bond * FZ_CS);
            const float4 cb = *reinterpret_cast<const float4*>(s_ctab + bond * FZ_CS + 4);
            const float c[FZ_K] = {ca.x * mult, ca.y * mult, ca.z * mult, ca.w * mult,
                                   cb.x * mult, cb.y * mult, cb.z * mult, cb.w * mult};
            const float4* hp = reinterpret_cast<const float4*>(&ws.h[src * FZ_HS + pass * MP]);
#pragma unroll
            for (int q = 0; q < MP / 4; ++q) {
              const float4 hv = hp[q];
              const float hs[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < FZ_K; ++k) acc[(q * 4 + j) * FZ_K + k] = fmaf(hs[j], c[k], acc[(q * 4 + j) * FZ_K + k]);
            }
          }
#pragma unroll
          for (int ch = 0; ch < MP * FZ_K / 64; ++ch) {
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = tc::pack2<FMT>(acc[ch * 64 + 2 * i], acc[ch * 64 + 2 * i + 1]);
            tc::tmem_st32(tZ + lane_off + (uint32_t)(pass * (MP * FZ_K / 2) + ch * 32), r);
          }
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, 128);
        // ------------------------------------------------------------ GEMM1: agg = Z . Wc
        if (t == 0) {
          tc::fence_after_thread_sync();
#pragma unroll
          for (int ks = 0; ks < D * FZ_K / 16; ++ks)
            tc::mma_f16_ts(tCagg, tZ + 8 * ks, tc::make_smem_desc(sw + ks * 1024, D * 16, 128), idesc32, ks > 0);
          tc::mma_commit(&ws.bar[0]);
        }
        tc::mbar_wait(&ws.bar[0], ph);
        tc::fence_after_thread_sync();
        {  // agg and h rows as 16-bit A operands (the Z columns are dead now)
          float v[32];
          tc::tmem_ld32(tCagg + lane_off, v);
          uint32_t r[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = tc::pack2<FMT>(v[2 * i], v[2 * i + 1]);
          tc::tmem_st16(tAagg + lane_off, r);
#pragma unroll
          for (int c = 0; c < D / 4; ++c) {
            const float4 x = reinterpret_cast<const float4*>(hrow)[c];
            r[2 * c] = tc::pack2<FMT>(x.x, x.y), r[2 * c + 1] = tc::pack2<FMT>(x.z, x.w);
          }
          tc::tmem_st16(tAh + lane_off, r);
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, 128);
        // ------------------------------------------------------------ GEMM2: [h | agg] . [Wz | Wr]; GEMM3a: agg . Wh[d:2d]
        if (t == 0) {
          tc::fence_after_thread_sync();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc::mma_f16_ts(tCzr, tAh + 8 * ks, tc::make_smem_desc(sw + FusedPack::OFF_BZR + ks * 2048, 2 * D * 16, 128),
                           idesc64, ks > 0);
          tc::mma_commit(&ws.bar[1]);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc::mma_f16_ts(tCht, tAagg + 8 * ks, tc::make_smem_desc(sw + FusedPack::OFF_BH + (ks + 2) * 1024, D * 16, 128),
                           idesc32, ks > 0);
        }
        tc::mbar_wait(&ws.bar[1], ph);
        tc::fence_after_thread_sync();
        float z[D];
        {
          float v[32];
          tc::tmem_ld32(tCzr + lane_off, v);
#pragma unroll
          for (int j = 0; j < D; ++j) z[j] = fz_sigmoid<PRECISE>(v[j] + bias[j]);
          tc::tmem_ld32(tCzr + D + lane_off, v);
          uint32_t r[16];
#pragma unroll
          for (int c = 0; c < D / 4; ++c) {
            const float4 x = reinterpret_cast<const float4*>(hrow)[c];
            const float r0 = fz_sigmoid<PRECISE>(v[4 * c] + bias[D + 4 * c]) * x.x;
            const float r1 = fz_sigmoid<PRECISE>(v[4 * c + 1] + bias[D + 4 * c + 1]) * x.y;
            const float r2 = fz_sigmoid<PRECISE>(v[4 * c + 2] + bias[D + 4 * c + 2]) * x.z;
            const float r3 = fz_sigmoid<PRECISE>(v[4 * c + 3] + bias[D + 4 * c + 3]) * x.w;
            r[2 * c] = tc::pack2<FMT>(r0, r1), r[2 * c + 1] = tc::pack2<FMT>(r2, r3);
          }
          tc::tmem_st16(tArh + lane_off, r);
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, 128);
        // ------------------------------------------------------------ GEMM3b: += (r*h) . Wh[0:d]
        if (t == 0) {
          tc::fence_after_thread_sync();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc::mma_f16_ts(tCht, tArh + 8 * ks, tc::make_smem_desc(sw + FusedPack::OFF_BH + ks * 1024, D * 16, 128), idesc32,
                           true);
          tc::mma_commit(&ws.bar[2]);
        }
        tc::mbar_wait(&ws.bar[2], ph);
        tc::fence_after_thread_sync();
        {  // candidate, blend, LayerNorm (biased variance, eps), residual  (models/layers.py:151-156)
          float gq[32], hq[32];
          tc::tmem_ld32(tCht + lane_off, gq);
#pragma unroll
          for (int c = 0; c < D / 4; ++c) {
            const float4 x = reinterpret_cast<const float4*>(hrow)[c];
            hq[4 * c] = x.x, hq[4 * c + 1] = x.y, hq[4 * c + 2] = x.z, hq[4 * c + 3] = x.w;
          }
          float mean = 0.f;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const float ht = fz_tanh<PRECISE>(gq[j] + bias[2 * D + j]);
            gq[j] = fmaf(z[j], ht - hq[j], hq[j]);
            mean += gq[j];
          }
          mean *= (1.0f / D);
          float var = 0.f;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const float cdev = gq[j] - mean;
            var = fmaf(cdev, cdev, var);
          }
          const float inv = PRECISE ? 1.0f / sqrtf(var * (1.0f / D) + a.eps) : rsqrtf(var * (1.0f / D) + a.eps);
#pragma unroll
          for (int c = 0; c < D / 4; ++c) {
            float4 o;
            o.x = fmaf((gq[4 * c] - mean) * inv, bias[3 * D + 4 * c], bias[4 * D + 4 * c]) + hq[4 * c];
            o.y = fmaf((gq[4 * c + 1] - mean) * inv, bias[3 * D + 4 * c + 1], bias[4 * D + 4 * c + 1]) + hq[4 * c + 1];
            o.z = fmaf((gq[4 * c + 2] - mean) * inv, bias[3 * D + 4 * c + 2], bias[4 * D + 4 * c + 2]) + hq[4 * c + 2];
            o.w = fmaf((gq[4 * c + 3] - mean) * inv, bias[3 * D + 4 * c + 3], bias[4 * D + 4 * c + 3]) + hq[4 * c + 3];
            reinterpret_cast<float4*>(hrow)[c] = o;
          }
        }
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, 128);
        ph ^= 1;
      }
      // ---------------------------------------------------------------- GlobalSumPool: warp per molecule, lane = column
      for (int mi = ms + (t >> 5); mi < me; mi += 4) {
        const int lo = ws.molp[mi] - a0, hi = min(ws.molp[mi + 1] - a0, FZ_ROWS);
        float sacc = 0.f;
        for (int r = lo; r < hi; ++r)
          if (ws.amask[r]) sacc += ws.h[r * FZ_HS + lane];
        a.pooled[(size_t)(base_mol + mi) * D + lane] = sacc;
      }
      tc::named_bar_sync(bar_id, 128);
      ms = me;
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

// =====================================================================================================================
// Second generation of the fused kernel, for IEEE-half operands ("h2"): same GEMM pipeline, cheaper and wider SIMT side.
//   * TWO threads per atom row (warps w and w+4 of a 256-thread "context" address the same TMEM lanes): each owns 16 of
//     the 32 state columns -- half of the Z row, half of every epilogue.  512 threads per SM (4 warps per
//     sub-partition) instead of 256 hide the tcgen05 round trips and the fixed-latency stalls that bounded the first
//     generation (profiles/r01_fused_v1: 42 % issue-slot use at 2 warps per sub-partition).
//   * Z rows are accumulated with packed HFMA2 (two products per lane-instruction) directly in the 16-bit pair layout
//     the tensor core reads, in ONE pass over the row's entries, so there is no fp32->fp16 pack and no second walk of
//     the entry list.  Neighbour states are gathered from a shared-memory copy of h that is already half precision and
//     lane-broadcast ((h_m, h_m) pairs); one rounding per accumulation step on a value that is rounded to half for the
//     MMA anyway.
//   * the fp32 state of a row lives in its owners' REGISTERS for all steps;
//   * rows are assigned to threads sorted by in-degree (counting sort with warp ballots), so that the 32 lanes of a
//     warp run the same number of entry iterations; the two contexts sort in opposite directions so that every SM
//     sub-partition gets light and heavy warps;
//   * LayerNorm statistics are exchanged between the two owners of a row through shared memory (sum, sum of squares);
//   * the index lines of the next tile are prefetched into L2 while the current tile computes.
constexpr int F2_CTX_THREADS = 256;

struct alignas(16) FusedWgSmem2 {
  uint32_t hb[FZ_ROWS * FZ_HS];  // half2 (h_m, h_m) per column; after the last step: fp32 h rows for the pooling
  float2 ln[2][FZ_ROWS];         // per half: (sum, sum of squares) of the blended row
  int molp[FZ_GROUP + 4];
  int se0[FZ_ROWS], se1[FZ_ROWS], said[FZ_ROWS];  // per natural row: entry range, atom id
  int cnt[4][8];
  unsigned char rowof[FZ_ROWS];
  unsigned char amask[FZ_ROWS];
  uint64_t bar[4];
};

__host__ __device__ inline int fused2_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack::BYTES + ctab + 2 * (int)sizeof(FusedWgSmem2) + (int)sizeof(FusedCtl);
}


template <bool PRECISE>
__global__ void __launch_bounds__(2 * F2_CTX_THREADS, 1) mpnn_fused_h2_kernel(const FusedArgs a) {
  constexpr int D = FZ_D, DH = FZ_D / 2;
  constexpr int FMT = tc::FMT_F16;
  constexpr int NT = 2 * F2_CTX_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 8, u = tid & 255;
  const int p = u & 127;   // row slot = TMEM lane
  const int hf = u >> 7;   // which 16 columns of the row this thread owns
  const int wq = warp & 3; // TMEM lane quarter == (p >> 5)
  const int wbytes = a.steps * FusedPack::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);  // per bond: (c0,c1) (c2,c3) (c4,c5) (c6,c7) as half2
  FusedWgSmem2& ws = reinterpret_cast<FusedWgSmem2*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + 2 * sizeof(FusedWgSmem2));

  const int tower = blockIdx.x >= a.n_cta_cat;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.packed + (size_t)tower * wbytes);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < wbytes / 16; i += NT) dst[i] = __ldg(src + i);
    for (int i = tid; i < a.bond_vocab; i += NT) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i);
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i + 1);
      s_ctab[i] = make_uint4(tc::pack_f16x2(c0.x, c0.y), tc::pack_f16x2(c0.z, c0.w), tc::pack_f16x2(c1.x, c1.y),
                             tc::pack_f16x2(c1.z, c1.w));
    }
  }
  if (warp == 0) tc::tmem_alloc<512>(&ctl.tmem_base);
  if (u == 0) {
    tc::mbar_init(&ws.bar[0], 1);
    tc::mbar_init(&ws.bar[1], 1);
    tc::mbar_init(&ws.bar[2], 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();

  const uint32_t sw0 = tc::smem_u32(smem);
  const uint32_t tbase = ctl.tmem_base + (uint32_t)(ctx * 256);
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  // TMEM columns of a context: [0,128) Z (K = 256), dead after GEMM1, then [0,16) h, [16,32) agg, [32,48) r*h operands;
  // [128,160) GEMM1 accumulator (agg); [160,224) z | r pre-activations; [224,256) candidate pre-activation.
  const uint32_t tZ = tbase, tAh = tbase, tAagg = tbase + 16, tArh = tbase + 32;
  const uint32_t tCagg = tbase + 128, tCzr = tbase + 160, tCht = tbase + 224;
  // B-operand descriptors of step 0; step s adds s * BYTES / 16 to the address field (no carry: smem < 256 KiB)
  const uint64_t dWc = tc::make_smem_desc(sw0, D * 16, 128);
  const uint64_t dBzr = tc::make_smem_desc(sw0 + FusedPack::OFF_BZR, 2 * D * 16, 128);
  const uint64_t dBh = tc::make_smem_desc(sw0 + FusedPack::OFF_BH, D * 16, 128);
  const bool mma_warp = (u >> 5) == 0;  // warp 0 of the context issues every MMA (one elected lane)
  const uint32_t idesc32 = tc::make_idesc(FMT, FZ_ROWS, D), idesc64 = tc::make_idesc(FMT, FZ_ROWS, 2 * D);
  const int bar_id = 1 + ctx;
  const bool descending = ctx & 1;

  const int P = a.n_pairs;
  const int n_groups = (P + FZ_GROUP - 1) / FZ_GROUP;
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  uint32_t ph = 0;
  FZ_PROF_DECL;

  for (int g = cta_in_tower * 2 + ctx; g < n_groups; g += n_cta_tower * 2) {
    const int m0 = g * FZ_GROUP, nm = min(FZ_GROUP, P - m0);
    const int base_mol = tower * P + m0;
    tc::named_bar_sync(bar_id, F2_CTX_THREADS);
    if (u <= nm) ws.molp[u] = __ldg(a.mol_ptr + base_mol + u);
    tc::named_bar_sync(bar_id, F2_CTX_THREADS);
    int ms = 0;
    while (ms < nm) {
      const int a0 = ws.molp[ms];
      int me = ms + 1;
      while (me < nm && ws.molp[me + 1] - a0 <= FZ_ROWS) ++me;
      int rows = ws.molp[me] - a0;
      if (rows > FZ_ROWS) {
        if (u == 0 && a.status) *a.status = 1;
        rows = FZ_ROWS;
      }
      // ---------------------------------------------------------------- natural row p: indices, in-degree key (warps 0-3)
      FZ_PROF_T(0);
      int key = 0, rank = 0;
      if (hf == 0) {
        const bool valid = p < rows;
        int aid = 0, e0 = 0, e1 = 0;
        if (valid) {
          aid = __ldg(a.atom_id + a0 + p);
          e0 = __ldg(a.row_ptr + a0 + p);
          e1 = __ldg(a.row_ptr + a0 + p + 1);
        }
        ws.se0[p] = e0, ws.se1[p] = e1, ws.said[p] = aid;
        ws.amask[p] = (valid && aid > 0) ? 1 : 0;  // models/layers.py:163
        key = min(e1 - e0, 7);
        int mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned m = __ballot_sync(0xffffffffu, key == k);
          if (lane == k) mine = __popc(m);
          if (key == k) rank = __popc(m & ((1u << lane) - 1u));
        }
        if (lane < 8) ws.cnt[wq][lane] = mine;
      }
      tc::named_bar_sync(bar_id, F2_CTX_THREADS);
      if (hf == 0) {
        int off = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c0 = ws.cnt[0][k], c1 = ws.cnt[1][k], c2 = ws.cnt[2][k], c3 = ws.cnt[3][k];
          if (k < key) off += c0 + c1 + c2 + c3;
          if (k == key) off += (wq > 0 ? c0 : 0) + (wq > 1 ? c1 : 0) + (wq > 2 ? c2 : 0);
        }
        const int slot = off + rank;
        ws.rowof[descending ? FZ_ROWS - 1 - slot : slot] = (unsigned char)p;
      }
      tc::named_bar_sync(bar_id, F2_CTX_THREADS);
      // ---------------------------------------------------------------- this thread owns columns [16 hf, 16 hf + 16) of row r
      const int r = ws.rowof[p];
      const int e0 = ws.se0[r], e1 = (FZ_DEBUG(a) & 1) ? e0 : ws.se1[r];
      uint32_t* hbrow = &ws.hb[r * FZ_HS + hf * DH];
      float h[DH];
      {  // Embedding(atom)
        const bool valid = r < rows;
        const int id = min(max(ws.said[r], 0), a.atom_vocab - 1);
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 x = valid ? __ldg(emb4 + id * (D / 4) + hf * (DH / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
          reinterpret_cast<uint4*>(hbrow)[c] =
              make_uint4(tc::pack_f16x2(x.x, x.x), tc::pack_f16x2(x.y, x.y), tc::pack_f16x2(x.z, x.z), tc::pack_f16x2(x.w, x.w));
        }
      }
      if (me < nm && hf == 0 && lane < 8) {  // index lines of the next tile -> L2 (it follows this tile in all arrays)
        const int an = ws.molp[me], en = ws.se1[max(rows, 1) - 1];
        if (wq == 0) prefetch_l2(a.atom_id + min(an + lane * 32, a.n_atoms - 1));
        if (wq == 1) prefetch_l2(a.row_ptr + min(an + lane * 32, a.n_atoms));
        if (wq == 2) prefetch_l2(a.col_src + min(en + lane * 32, a.n_unique - 1));
        if (wq == 3) prefetch_l2(a.edge_bm + min(en + lane * 32, a.n_unique - 1));
      }
      tc::named_bar_sync(bar_id, F2_CTX_THREADS);

      FZ_PROF_T(1);
      for (int s = 0; s < a.steps; ++s) {
        FZ_PROF_T(2);
        const uint64_t dstep = (uint64_t)(s * (FusedPack::BYTES / 16));
        const float* bias = reinterpret_cast<const float*>(smem + s * FusedPack::BYTES + FusedPack::OFF_BIAS) + hf * DH;
        // ------------------------------------------------------------ half of the Z row (half2 accumulators) -> TMEM
        {
          __half2 acc[DH * FZ_K / 2];
#pragma unroll
          for (int i = 0; i < DH * FZ_K / 2; ++i) acc[i] = __half2(__ushort_as_half(0), __ushort_as_half(0));
#pragma unroll 1
          for (int e = e0; e < e1; ++e) {
            const int bm = __ldg(a.edge_bm + e);
            int src = __ldg(a.col_src + e) - a0;
            src = min(max(src, 0), FZ_ROWS - 1);
            const __half2 mult = __float2half2_rn((float)(bm >> 16));
            const int bond = min(bm & 0xffff, a.bond_vocab - 1);
            const uint4 cu = s_ctab[bond];
            __half2 c[4];
            c[0] = __hmul2(*reinterpret_cast<const __half2*>(&cu.x), mult);
            c[1] = __hmul2(*reinterpret_cast<const __half2*>(&cu.y), mult);
            c[2] = __hmul2(*reinterpret_cast<const __half2*>(&cu.z), mult);
            c[3] = __hmul2(*reinterpret_cast<const __half2*>(&cu.w), mult);
            const uint4* hp = reinterpret_cast<const uint4*>(&ws.hb[src * FZ_HS + hf * DH]);
#pragma unroll
            for (int q = 0; q < DH / 4; ++q) {
              const uint4 hv = hp[q];
              const __half2 hm[4] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y),
                                     *reinterpret_cast<const __half2*>(&hv.z), *reinterpret_cast<const __half2*>(&hv.w)};
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[(4 * q + i) * 4 + j] = __hfma2(hm[i], c[j], acc[(4 * q + i) * 4 + j]);
            }
          }
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t rr[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) rr[i] = *reinterpret_cast<const uint32_t*>(&acc[ch * 32 + i]);
            tc::tmem_st32(tZ + lane_off + (uint32_t)(hf * 64 + ch * 32), rr);
          }
        }
        tc::tmem_wait_st();
        FZ_PROF_T(3);
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F2_CTX_THREADS);
        FZ_PROF_T(4);
        // ------------------------------------------------------------ GEMM1: agg = Z . Wc
        if (mma_warp && !(FZ_DEBUG(a) & 2)) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < D * FZ_K / 16; ++ks)
              tc::mma_f16_ts(tCagg, tZ + 8 * ks, dWc + dstep + (uint64_t)(ks * 64), idesc32, ks > 0);
            tc::mma_commit(&ws.bar[0]);
          }
          __syncwarp();
        }
        if (!(FZ_DEBUG(a) & 2)) tc::mbar_wait(&ws.bar[0], ph);
        tc::fence_after_thread_sync();
        FZ_PROF_T(5);
        {
          float v[16];
          tc::tmem_ld16(tCagg + lane_off + hf * DH, v);
          uint32_t rr[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rr[i] = tc::pack_f16x2(v[2 * i], v[2 * i + 1]);
          tc::tmem_st8(tAagg + lane_off + hf * 8, rr);
#pragma unroll
          for (int i = 0; i < 8; ++i) rr[i] = tc::pack_f16x2(h[2 * i], h[2 * i + 1]);
          tc::tmem_st8(tAh + lane_off + hf * 8, rr);
        }
        tc::tmem_wait_st();
        FZ_PROF_T(6);
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F2_CTX_THREADS);
        FZ_PROF_T(7);
        // ------------------------------------------------------------ GEMM2 / GEMM3a
        if (mma_warp && !(FZ_DEBUG(a) & 2)) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma_f16_ts(tCzr, tAh + 8 * ks, dBzr + dstep + (uint64_t)(ks * 128), idesc64, ks > 0);
            tc::mma_commit(&ws.bar[1]);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::mma_f16_ts(tCht, tAagg + 8 * ks, dBh + dstep + (uint64_t)((ks + 2) * 64), idesc32, ks > 0);
          }
          __syncwarp();
        }
        if (!(FZ_DEBUG(a) & 2)) tc::mbar_wait(&ws.bar[1], ph);
        tc::fence_after_thread_sync();
        FZ_PROF_T(8);
        float z[DH];
        {
          float v[16];
          tc::tmem_ld16(tCzr + lane_off + hf * DH, v);
#pragma unroll
          for (int j = 0; j < DH; ++j) z[j] = fz_sigmoid<PRECISE>(v[j] + bias[j]);
          tc::tmem_ld16(tCzr + D + lane_off + hf * DH, v);
          uint32_t rr[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float r0 = fz_sigmoid<PRECISE>(v[2 * i] + bias[D + 2 * i]) * h[2 * i];
            const float r1 = fz_sigmoid<PRECISE>(v[2 * i + 1] + bias[D + 2 * i + 1]) * h[2 * i + 1];
            rr[i] = tc::pack_f16x2(r0, r1);
          }
          tc::tmem_st8(tArh + lane_off + hf * 8, rr);
        }
        tc::tmem_wait_st();
        FZ_PROF_T(9);
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F2_CTX_THREADS);
        FZ_PROF_T(10);
        // ------------------------------------------------------------ GEMM3b
        if (mma_warp && !(FZ_DEBUG(a) & 2)) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCht, tArh + 8 * ks, dBh + dstep + (uint64_t)(ks * 64), idesc32, true);
            tc::mma_commit(&ws.bar[2]);
          }
          __syncwarp();
        }
        if (!(FZ_DEBUG(a) & 2)) tc::mbar_wait(&ws.bar[2], ph);
        tc::fence_after_thread_sync();
        FZ_PROF_T(11);
        {  // candidate, blend, LayerNorm, residual  (models/layers.py:151-156)
          float gq[16];
          tc::tmem_ld16(tCht + lane_off + hf * DH, gq);
          float sum = 0.f, sq = 0.f;
#pragma unroll
          for (int j = 0; j < DH; ++j) {
            const float ht = fz_tanh<PRECISE>(gq[j] + bias[2 * D + j]);
            gq[j] = fmaf(z[j], ht - h[j], h[j]);
            sum += gq[j];
            sq = fmaf(gq[j], gq[j], sq);
          }
          ws.ln[hf][p] = make_float2(sum, sq);
          FZ_PROF_T(12);
          tc::fence_before_thread_sync();
          tc::named_bar_sync(bar_id, F2_CTX_THREADS);
          FZ_PROF_T(13);
          const float2 other = ws.ln[hf ^ 1][p];
          const float mean = (sum + other.x) * (1.0f / D);
          const float var = fmaxf((sq + other.y) * (1.0f / D) - mean * mean, 0.f);  // biased variance
          const float inv = PRECISE ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
#pragma unroll
          for (int j = 0; j < DH; ++j) h[j] = fmaf((gq[j] - mean) * inv, bias[3 * D + j], bias[4 * D + j]) + h[j];
          if (s + 1 < a.steps) {
#pragma unroll
            for (int c = 0; c < DH / 4; ++c)
              reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(tc::pack_f16x2(h[4 * c], h[4 * c]), tc::pack_f16x2(h[4 * c + 1], h[4 * c + 1]),
                                                              tc::pack_f16x2(h[4 * c + 2], h[4 * c + 2]), tc::pack_f16x2(h[4 * c + 3], h[4 * c + 3]));
          } else {  // last step: fp32 rows for the pooling (every gather of this tile is done)
#pragma unroll
            for (int c = 0; c < DH / 4; ++c)
              reinterpret_cast<float4*>(hbrow)[c] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
          }
        }
        FZ_PROF_T(14);
        tc::named_bar_sync(bar_id, F2_CTX_THREADS);
        FZ_PROF_T(15);
        ph ^= 1;
      }
      // ---------------------------------------------------------------- GlobalSumPool: warp per molecule, lane = column
      {
        const float* hfp = reinterpret_cast<const float*>(ws.hb);
        for (int mi = ms + (u >> 5); mi < me; mi += F2_CTX_THREADS / 32) {
          const int lo = ws.molp[mi] - a0, hi = min(ws.molp[mi + 1] - a0, FZ_ROWS);
          float sacc = 0.f;
          for (int rr = lo; rr < hi; ++rr)
            if (ws.amask[rr]) sacc += hfp[rr * FZ_HS + lane];
          a.pooled[(size_t)(base_mol + mi) * D + lane] = sacc;
        }
      }
      FZ_PROF_T(16);
      tc::named_bar_sync(bar_id, F2_CTX_THREADS);
      FZ_PROF_T(17);
      ms = me;
    }
  }
  FZ_PROF_FLUSH;
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

// =====================================================================================================================
// Third generation ("h2x", the default for half operands): NCTX independent 128-thread contexts per CTA, one thread per
// atom row, 128 TMEM columns per context.  The second generation is bounded by latency: a context is a serial chain of
// barriers and tcgen05 round trips, and TMEM (2 x 256 columns) allowed only two of them per SM.  Here
//   * Z is built and multiplied in two K halves (bond-embedding components k < 4, then k >= 4) through the SAME 64
//     columns -- 64 half2 accumulators per thread -- and GEMM1 of the first half runs under the build of the second;
//   * accumulators are recycled: [64,96) agg, then [64,128) z|r, then [64,96) candidate; operands h / agg / r*h reuse
//     the Z columns; the candidate GEMM is one K = 64 chain [agg | r*h] . Wh issued after the gate epilogue;
// so a context needs 128 columns and four of them fit (512 threads, <= 128 registers each; three at <= 168 registers).
// Row ownership, degree sort, packed-HFMA2 Z build, register-resident fp32 state, warp-uniform MMA issue: as above.
constexpr int F3_CTX_THREADS = 128;
constexpr int F3_ECAP = 688;  // decoded entries staged per tile (tiles with more entries read the global arrays)

struct alignas(16) FusedWgSmem3 {
  uint32_t hb[FZ_ROWS * FZ_HS];  // words 0..15 of a row: h as packed halves; after the last step: fp32 h rows for the pooling
  int molp[FZ_GROUP + 4];
  int se0[FZ_ROWS], se1[FZ_ROWS], said[FZ_ROWS];
  int cnt[4][8];
  int mole[FZ_GROUP + 4];  // compact feed: first entry of every molecule of the group
  int wsum[4];
  uint32_t ent[F3_ECAP];  // decoded entries of the tile: src row | bond << 8 | mult (half bits) << 16
  unsigned char rowof[FZ_ROWS];
  unsigned char amask[FZ_ROWS];
  uint64_t bar[4];
};

__host__ __device__ inline int fused3_smem_bytes(int steps, int bond_vocab, int nctx) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack::BYTES + ctab + nctx * (int)sizeof(FusedWgSmem3) + (int)sizeof(FusedCtl);
}

template <bool PRECISE, int NCTX, bool COMPACT>
__global__ void __launch_bounds__(NCTX * F3_CTX_THREADS, 1) mpnn_fused_h2x_kernel(const FusedArgs a) {
  constexpr int D = FZ_D;
  constexpr int FMT = tc::FMT_F16;
  constexpr int NT = NCTX * F3_CTX_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 7, t = tid & 127, wq = warp & 3;
  const int wbytes = a.steps * FusedPack::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);
  FusedWgSmem3& ws = reinterpret_cast<FusedWgSmem3*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + NCTX * sizeof(FusedWgSmem3));

  const int tower = blockIdx.x >= a.n_cta_cat;
  // Resident weights of this tower (all steps, 127 KB): ONE TMA bulk copy, issued by one thread and
  // counted in bytes on an mbarrier; the bond-coefficient table is converted by the threads meanwhile.
  if (tid == 0) {
    tc::mbar_init(&ctl.wbar, 1);
    tc::mbar_fence_init();
    tc::mbar_arrive_expect_tx(&ctl.wbar, (uint32_t)wbytes);
    tc::bulk_copy_g2s(smem, a.packed + (size_t)tower * wbytes, (uint32_t)wbytes, &ctl.wbar);
  }
  for (int i = tid; i < a.bond_vocab; i += NT) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i + 1);
    s_ctab[i] = make_uint4(tc::pack_f16x2(c0.x, c0.y), tc::pack_f16x2(c0.z, c0.w), tc::pack_f16x2(c1.x, c1.y),
                           tc::pack_f16x2(c1.z, c1.w));
  }
  if (warp == 0) tc::tmem_alloc<512>(&ctl.tmem_base);
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::mbar_init(&ws.bar[i], 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  tc::mbar_wait(&ctl.wbar, 0);  // weights have landed (async proxy writes: visible to tcgen05.mma without a proxy fence)

  const uint32_t sw0 = tc::smem_u32(smem);
  const uint32_t tbase = ctl.tmem_base + (uint32_t)(ctx * 128);
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t tZ = tbase, tAh = tbase, tAagg = tbase + 16, tArh = tbase + 32, tOnes = tbase + 48;
  const uint32_t tCagg = tbase + 64, tCzr = tbase + 64, tCht = tbase + 64;
  const uint32_t idesc32 = tc::make_idesc(FMT, FZ_ROWS, D), idesc64 = tc::make_idesc(FMT, FZ_ROWS, 2 * D);
  const uint64_t dWc = tc::make_smem_desc(sw0, D * 16, 128);
  const uint64_t dBzr = tc::make_smem_desc(sw0 + FusedPack::OFF_BZR, 2 * D * 16, 128);
  const uint64_t dBh = tc::make_smem_desc(sw0 + FusedPack::OFF_BH, D * 16, 128);
  const uint64_t dBBzr = tc::make_smem_desc(sw0 + FusedPack::OFF_BBZR, 2 * D * 16, 128);
  const uint64_t dBBh = tc::make_smem_desc(sw0 + FusedPack::OFF_BBH, D * 16, 128);
  const bool mma_warp = (t >> 5) == 0;
  const int bar_id = 1 + ctx;
  const bool descending = ctx & 1;

  const int P = a.n_pairs;
  const int n_groups = (P + FZ_GROUP - 1) / FZ_GROUP;
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  uint32_t ph = 0;
  [[maybe_unused]] const int u = t;
  FZ_PROF_DECL;

  for (int g = cta_in_tower * NCTX + ctx; g < n_groups; g += n_cta_tower * NCTX) {
    const int m0 = g * FZ_GROUP, nm = min(FZ_GROUP, P - m0);
    const int base_mol = tower * P + m0;
    tc::named_bar_sync(bar_id, F3_CTX_THREADS);
    if (t <= nm) {
      ws.molp[t] = __ldg(a.mol_ptr + base_mol + t);
      if (COMPACT) ws.mole[t] = __ldg(a.mol_eptr + base_mol + t);
    }
    tc::named_bar_sync(bar_id, F3_CTX_THREADS);
    int ms = 0;
    while (ms < nm) {
      const int a0 = ws.molp[ms];
      int me = ms + 1;
      while (me < nm && ws.molp[me + 1] - a0 <= FZ_ROWS) ++me;
      int rows = ws.molp[me] - a0;
      if (rows > FZ_ROWS) {
        if (t == 0 && a.status) *a.status = 1;
        rows = FZ_ROWS;
      }
      FZ_PROF_T(0);
      // ---------------------------------------------------------------- natural row t: indices, in-degree key
      int key, rank = 0;
      {
        const bool valid = t < rows;
        int aid = 0, e0 = 0, e1 = 0;
        if (!COMPACT) {
          if (valid) {
            aid = __ldg(a.atom_id + a0 + t);
            e0 = __ldg(a.row_ptr + a0 + t);
            e1 = __ldg(a.row_ptr + a0 + t + 1);
          }
        } else {  // row_ptr of the tile = first entry of its first molecule + exclusive scan of the in-degrees
          const int aw = valid ? (int)__ldg(a.atom_w + a0 + t) : 0;
          aid = aw & 0xff;
          const int deg = aw >> 8;
          int incl = deg;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
          }
          if (lane == 31) ws.wsum[wq] = incl;
          tc::named_bar_sync(bar_id, F3_CTX_THREADS);
          int base = ws.mole[ms];
          for (int w = 0; w < wq; ++w) base += ws.wsum[w];
          e1 = base + incl;
          e0 = e1 - deg;
        }
        ws.se0[t] = e0, ws.se1[t] = e1, ws.said[t] = aid;
        ws.amask[t] = (valid && aid > 0) ? 1 : 0;  // models/layers.py:163
        key = min(e1 - e0, 7);
        int mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned m = __ballot_sync(0xffffffffu, key == k);
          if (lane == k) mine = __popc(m);
          if (key == k) rank = __popc(m & ((1u << lane) - 1u));
        }
        if (lane < 8) ws.cnt[wq][lane] = mine;
      }
      FZ_PROF_T(1);
      tc::named_bar_sync(bar_id, F3_CTX_THREADS);
      FZ_PROF_T(2);
      {
        int off = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c0 = ws.cnt[0][k], c1 = ws.cnt[1][k], c2 = ws.cnt[2][k], c3 = ws.cnt[3][k];
          if (k < key) off += c0 + c1 + c2 + c3;
          if (k == key) off += (wq > 0 ? c0 : 0) + (wq > 1 ? c1 : 0) + (wq > 2 ? c2 : 0);
        }
        const int slot = off + rank;
        ws.rowof[descending ? FZ_ROWS - 1 - slot : slot] = (unsigned char)t;
      }
      FZ_PROF_T(3);
      tc::named_bar_sync(bar_id, F3_CTX_THREADS);
      FZ_PROF_T(4);
      // ---------------------------------------------------------------- decoded entries of the tile -> shared memory
      const int E0 = ws.se0[0], n_ent = ws.se1[max(rows, 1) - 1] - E0;
      const bool staged = n_ent <= F3_ECAP && a.bond_vocab <= 256;
      if (staged)
        for (int i = t; i < n_ent; i += F3_CTX_THREADS) {
          if (COMPACT) {  // molecule-local source -> tile row: the entry's molecule is found from the entry offsets
            const uint32_t w = __ldg(a.edge_w + E0 + i);
            int mrow = 0;
            for (int mi = ms; mi < me; ++mi)
              if (ws.mole[mi] <= E0 + i) mrow = ws.molp[mi] - a0;
            const uint32_t src = (uint32_t)min((int)(w & 0xffu) + mrow, FZ_ROWS - 1);
            const uint32_t bond = min((w >> 8) & 0xffu, (uint32_t)(a.bond_vocab - 1));
            ws.ent[i] = src | (bond << 8) | ((uint32_t)__half_as_ushort(__float2half_rn((float)((w >> 16) & 0xffu))) << 16);
            continue;
          }
          {
            const int bm = __ldg(a.edge_bm + E0 + i);
            const int src = min(max(__ldg(a.col_src + E0 + i) - a0, 0), FZ_ROWS - 1);
            const int bond = min(bm & 0xffff, a.bond_vocab - 1);
            ws.ent[i] = (uint32_t)src | ((uint32_t)bond << 8) | ((uint32_t)__half_as_ushort(__float2half_rn((float)(bm >> 16))) << 16);
          }
        }
      FZ_PROF_T(5);
      // ---------------------------------------------------------------- thread t owns row r
      const int r = ws.rowof[t];
      const int e0 = ws.se0[r], e1 = (FZ_DEBUG(a) & 1) ? e0 : ws.se1[r];
      uint32_t* hbrow = &ws.hb[r * FZ_HS];
      float h[D];
      {  // Embedding(atom)
        const bool valid = r < rows;
        const int id = min(max(ws.said[r], 0), a.atom_vocab - 1);
#pragma unroll
        for (int c = 0; c < D / 4; ++c) {
          const float4 x = valid ? __ldg(emb4 + id * (D / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
          reinterpret_cast<uint2*>(hbrow)[c] = make_uint2(tc::pack_f16x2(x.x, x.y), tc::pack_f16x2(x.z, x.w));
        }
      }
      if (me < nm && lane < 8) {  // index lines of the next tile -> L2
        const int an = ws.molp[me], en = ws.se1[max(rows, 1) - 1];
        if (!COMPACT) {
          if (wq == 0) prefetch_l2(a.atom_id + min(an + lane * 32, a.n_atoms - 1));
          if (wq == 1) prefetch_l2(a.row_ptr + min(an + lane * 32, a.n_atoms));
          if (wq == 2) prefetch_l2(a.col_src + min(en + lane * 32, a.n_unique - 1));
          if (wq == 3) prefetch_l2(a.edge_bm + min(en + lane * 32, a.n_unique - 1));
        } else {
          if (wq == 0 && lane < 2) prefetch_l2(a.atom_w + min(an + lane * 64, a.n_atoms - 1));
          if (wq == 2) prefetch_l2(a.edge_w + min(en + lane * 32, a.n_unique - 1));
        }
      }
      FZ_PROF_T(6);
      tc::named_bar_sync(bar_id, F3_CTX_THREADS);
      FZ_PROF_T(7);

      for (int s = 0; s < a.steps; ++s) {
        const uint64_t dstep = (uint64_t)(s * (FusedPack::BYTES / 16));
        const float* bias = reinterpret_cast<const float*>(smem + s * FusedPack::BYTES + FusedPack::OFF_BIAS);
        // ------------------------------------------------------------ Z in two K halves -> TMEM -> GEMM1
#pragma unroll 1
        for (int hz = 0; hz < 2; ++hz) {
          FZ_PROF_T(8);
          __half2 acc[D * 2];
#pragma unroll
          for (int i = 0; i < D * 2; ++i) acc[i] = __half2(__ushort_as_half(0), __ushort_as_half(0));
          if (staged) {
            uint32_t en = e0 < e1 ? ws.ent[e0 - E0] : 0u;
#pragma unroll 1
            for (int e = e0; e < e1; ++e) {
              const uint32_t ec = en;
              if (e + 1 < e1) en = ws.ent[e + 1 - E0];  // next entry's descriptor is in flight during this one's FMAs
              const uint2 cu = reinterpret_cast<const uint2*>(s_ctab + ((ec >> 8) & 0xff))[hz];
              const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
              const __half2 mult = *reinterpret_cast<const __half2*>(&mbits);
              const __half2 c0 = __hmul2(*reinterpret_cast<const __half2*>(&cu.x), mult);
              const __half2 c1 = __hmul2(*reinterpret_cast<const __half2*>(&cu.y), mult);
              const uint4* hp = reinterpret_cast<const uint4*>(&ws.hb[(ec & 0xff) * FZ_HS]);
#pragma unroll
              for (int q = 0; q < D / 8; ++q) {  // 8 columns per 16-byte read; HFMA2 broadcasts the low / high half
                const uint4 hv = hp[q];
                const __half2 hw[4] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y),
                                       *reinterpret_cast<const __half2*>(&hv.z), *reinterpret_cast<const __half2*>(&hv.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __half2 lo = __low2half2(hw[i]), hi = __high2half2(hw[i]);
                  const int m = 8 * q + 2 * i;
                  acc[m * 2] = __hfma2(lo, c0, acc[m * 2]);
                  acc[m * 2 + 1] = __hfma2(lo, c1, acc[m * 2 + 1]);
                  acc[m * 2 + 2] = __hfma2(hi, c0, acc[m * 2 + 2]);
                  acc[m * 2 + 3] = __hfma2(hi, c1, acc[m * 2 + 3]);
                }
              }
            }
          } else {
            int mbase = 0;  // COMPACT: first row of the molecule that owns row r (recomputed here: rare path, no live register)
            if (COMPACT)
              for (int mi = ms; mi < me; ++mi)
                if (ws.molp[mi] - a0 <= r) mbase = ws.molp[mi] - a0;
#pragma unroll 1
            for (int e = e0; e < e1; ++e) {
              int bm, src;
              if (!COMPACT) {
                bm = __ldg(a.edge_bm + e);
                src = __ldg(a.col_src + e) - a0;
              } else {
                const uint32_t w = __ldg(a.edge_w + e);
                bm = (int)(((w >> 8) & 0xffu) | ((w >> 16) & 0xffu) << 16);
                src = (int)(w & 0xffu) + mbase;
              }
              src = min(max(src, 0), FZ_ROWS - 1);
              const __half2 mult = __float2half2_rn((float)(bm >> 16));
              const int bond = min(bm & 0xffff, a.bond_vocab - 1);
              const uint2 cu = reinterpret_cast<const uint2*>(s_ctab + bond)[hz];
              const __half2 c0 = __hmul2(*reinterpret_cast<const __half2*>(&cu.x), mult);
              const __half2 c1 = __hmul2(*reinterpret_cast<const __half2*>(&cu.y), mult);
              const uint4* hp = reinterpret_cast<const uint4*>(&ws.hb[src * FZ_HS]);
#pragma unroll
              for (int q = 0; q < D / 8; ++q) {  // 8 columns per 16-byte read; HFMA2 broadcasts the low / high half
                const uint4 hv = hp[q];
                const __half2 hw[4] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y),
                                       *reinterpret_cast<const __half2*>(&hv.z), *reinterpret_cast<const __half2*>(&hv.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __half2 lo = __low2half2(hw[i]), hi = __high2half2(hw[i]);
                  const int m = 8 * q + 2 * i;
                  acc[m * 2] = __hfma2(lo, c0, acc[m * 2]);
                  acc[m * 2 + 1] = __hfma2(lo, c1, acc[m * 2 + 1]);
                  acc[m * 2 + 2] = __hfma2(hi, c0, acc[m * 2 + 2]);
                  acc[m * 2 + 3] = __hfma2(hi, c1, acc[m * 2 + 3]);
                }
              }
            }
          }
          FZ_PROF_T(11 + hz);  // Z half built
          if (hz == 1 && !(FZ_DEBUG(a) & 2)) {  // GEMM1a must have consumed the first half before its columns are rewritten
            tc::mbar_wait(&ws.bar[3], ph);
            tc::fence_after_thread_sync();
          }
          FZ_PROF_T(13);  // wait for GEMM1a (second half only)
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t rr[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) rr[i] = *reinterpret_cast<const uint32_t*>(&acc[ch * 32 + i]);
            tc::tmem_st32(tZ + lane_off + (uint32_t)(ch * 32), rr);
          }
          tc::tmem_wait_st();
          FZ_PROF_T(14);  // Z half stored
          tc::fence_before_thread_sync();
          tc::named_bar_sync(bar_id, F3_CTX_THREADS);
          FZ_PROF_T(15 + hz);  // barrier after the Z half
          if (mma_warp && !(FZ_DEBUG(a) & 2)) {
            tc::fence_after_thread_sync();
            if (tc::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc::mma_f16_ts(tCagg, tZ + 8 * ks, dWc + dstep + (uint64_t)(hz * (FusedPack::WC_BYTES / 32) + ks * 64), idesc32,
                               hz > 0 || ks > 0);
              tc::mma_commit(hz == 0 ? &ws.bar[3] : &ws.bar[0]);
            }
            __syncwarp();
          }
        }
        FZ_PROF_T(17);  // MMA issue of the second half
        if (!(FZ_DEBUG(a) & 2)) tc::mbar_wait(&ws.bar[0], ph);
        tc::fence_after_thread_sync();
        FZ_PROF_T(18);  // wait for GEMM1
        {  // agg and h as 16-bit A operands
          float v[32];
          tc::tmem_ld32(tCagg + lane_off, v);
          uint32_t rr[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) rr[i] = tc::pack_f16x2(v[2 * i], v[2 * i + 1]);
          tc::tmem_st16(tAagg + lane_off, rr);
#pragma unroll
          for (int i = 0; i < 16; ++i) rr[i] = tc::pack_f16x2(h[2 * i], h[2 * i + 1]);
          tc::tmem_st16(tAh + lane_off, rr);
          const uint32_t ones[8] = {0x00003c00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};  // (1, 0, ..., 0): the bias K-step
          tc::tmem_st8(tOnes + lane_off, ones);
        }
        tc::tmem_wait_st();
        FZ_PROF_T(19);  // operands written
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F3_CTX_THREADS);
        FZ_PROF_T(20);  // barrier before GEMM2
        // ------------------------------------------------------------ GEMM2: 0.5 ([h | agg | 1] . [Wz | Wr ; bz | br])
        if (mma_warp && !(FZ_DEBUG(a) & 2)) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma_f16_ts(tCzr, tAh + 8 * ks, dBzr + dstep + (uint64_t)(ks * 128), idesc64, ks > 0);
            tc::mma_f16_ts(tCzr, tOnes, dBBzr + dstep, idesc64, true);
            tc::mma_commit(&ws.bar[1]);
          }
          __syncwarp();
        }
        if (!(FZ_DEBUG(a) & 2)) tc::mbar_wait(&ws.bar[1], ph);
        tc::fence_after_thread_sync();
        FZ_PROF_T(21);  // wait for GEMM2
        float z[D];
        {
          float v[32];
          tc::tmem_ld32(tCzr + lane_off, v);
#pragma unroll
          for (int j = 0; j < D; ++j) z[j] = fz_sigmoid_half<PRECISE>(v[j]);
          tc::tmem_ld32(tCzr + D + lane_off, v);
          uint32_t rr[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float r0 = fz_sigmoid_half<PRECISE>(v[2 * i]) * h[2 * i];
            const float r1 = fz_sigmoid_half<PRECISE>(v[2 * i + 1]) * h[2 * i + 1];
            rr[i] = tc::pack_f16x2(r0, r1);
          }
          tc::tmem_st16(tArh + lane_off, rr);
        }
        tc::tmem_wait_st();
        FZ_PROF_T(22);  // gates
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F3_CTX_THREADS);
        FZ_PROF_T(23);  // barrier before GEMM3
        // ------------------------------------------------------------ GEMM3: [agg | r*h] . [Wh[d:2d] ; Wh[0:d]]
        if (mma_warp && !(FZ_DEBUG(a) & 2)) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_f16_ts(tCht, tAagg + 8 * ks, dBh + dstep + (uint64_t)((ks < 2 ? ks + 2 : ks - 2) * 64), idesc32, ks > 0);
            tc::mma_f16_ts(tCht, tOnes, dBBh + dstep, idesc32, true);
            tc::mma_commit(&ws.bar[2]);
          }
          __syncwarp();
        }
        if (!(FZ_DEBUG(a) & 2)) tc::mbar_wait(&ws.bar[2], ph);
        tc::fence_after_thread_sync();
        FZ_PROF_T(24);  // wait for GEMM3
        {  // candidate, blend, LayerNorm (biased variance, eps), residual  (models/layers.py:151-156)
          float gq[32];
          tc::tmem_ld32(tCht + lane_off, gq);
          float mean = 0.f, sq = 0.f;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const float ht = fz_tanh<PRECISE>(gq[j]);
            gq[j] = fmaf(z[j], ht - h[j], h[j]);
            mean += gq[j];
            sq = fmaf(gq[j], gq[j], sq);
          }
          mean *= (1.0f / D);
          const float var = fmaxf(fmaf(sq, 1.0f / D, -mean * mean), 0.f);  // biased variance
          const float inv = PRECISE ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
#pragma unroll
          for (int j = 0; j < D; ++j) h[j] = fmaf((gq[j] - mean) * inv, bias[3 * D + j], bias[4 * D + j]) + h[j];
          if (s + 1 < a.steps) {
#pragma unroll
            for (int c = 0; c < D / 8; ++c)
              reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(tc::pack_f16x2(h[8 * c], h[8 * c + 1]), tc::pack_f16x2(h[8 * c + 2], h[8 * c + 3]),
                                                              tc::pack_f16x2(h[8 * c + 4], h[8 * c + 5]), tc::pack_f16x2(h[8 * c + 6], h[8 * c + 7]));
          } else {
#pragma unroll
            for (int c = 0; c < D / 4; ++c)
              reinterpret_cast<float4*>(hbrow)[c] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
          }
        }
        FZ_PROF_T(25);  // candidate, blend, LayerNorm
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F3_CTX_THREADS);
        FZ_PROF_T(26);  // barrier at the end of the step
        ph ^= 1;
      }
      FZ_PROF_T(27);
      // ---------------------------------------------------------------- GlobalSumPool
      {
        const float* hfp = reinterpret_cast<const float*>(ws.hb);
        for (int mi = ms + (t >> 5); mi < me; mi += 4) {
          const int lo = ws.molp[mi] - a0, hi = min(ws.molp[mi + 1] - a0, FZ_ROWS);
          // four interleaved partial sums (rows lo+0, lo+4, ... etc.), combined in a fixed order: the row loop is a
          // chain of dependent shared-memory loads and adds, and it sits on every tile's critical path
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int rr = lo;
          for (; rr + 4 <= hi; rr += 4) {
            const float v0 = hfp[rr * FZ_HS + lane], v1 = hfp[(rr + 1) * FZ_HS + lane];
            const float v2 = hfp[(rr + 2) * FZ_HS + lane], v3 = hfp[(rr + 3) * FZ_HS + lane];
            const uchar4 mk = make_uchar4(ws.amask[rr], ws.amask[rr + 1], ws.amask[rr + 2], ws.amask[rr + 3]);
            s0 += mk.x ? v0 : 0.f, s1 += mk.y ? v1 : 0.f, s2 += mk.z ? v2 : 0.f, s3 += mk.w ? v3 : 0.f;
          }
          for (; rr < hi; ++rr) s0 += ws.amask[rr] ? hfp[rr * FZ_HS + lane] : 0.f;
          a.pooled[(size_t)(base_mol + mi) * D + lane] = (s0 + s1) + (s2 + s3);
        }
      }
      FZ_PROF_T(9);
      tc::named_bar_sync(bar_id, F3_CTX_THREADS);
      FZ_PROF_T(10);
      ms = me;
    }
  }
  FZ_PROF_FLUSH;
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

}  // namespace imp

using namespace imp;

extern "C" int64_t imp_fused_pack_bytes(int32_t d, int32_t bond_dim) {
  return (d == FZ_D && bond_dim == FZ_K) ? (int64_t)FusedPack::BYTES : (int64_t)IMP_ERR_DIM;
}

extern "C" int imp_fused_pack(const float* d_bond_transform, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                              int32_t flags, void* d_packed, void* stream) {
  IMP_REQUIRE(d_bond_transform && d_packed && w && w->Wz && w->bz && w->Wr && w->br && w->Wh && w->bh && w->gamma && w->beta,
              IMP_ERR_ARG, "imp_fused_pack: null pointer");
  IMP_REQUIRE(d == FZ_D && bond_dim == FZ_K, IMP_ERR_DIM, "imp_fused_pack: the fused path is built for atom_dim %d, bond_dim %d (got %d, %d)",
              FZ_D, FZ_K, d, bond_dim);
  const int n = FZ_D * FZ_D * FZ_K;
  const bool khalf = (flags & IMP_TC_FP16) && !(flags & (IMP_TC_F32_ZBUILD | IMP_TC_TWO_THREADS_PER_ROW));
  const bool gen4 = khalf && (flags & IMP_TC_GEN4) && !(flags & (IMP_TC_GEN3 | IMP_TC_THREE_CONTEXTS));
  if (khalf)
    fused_pack_kernel<tc::FMT_F16, true><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_bond_transform, *w, (unsigned char*)d_packed,
                                                                                           gen4 ? 0.5f : 1.0f);
  else if (flags & IMP_TC_FP16)
    fused_pack_kernel<tc::FMT_F16, false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_bond_transform, *w, (unsigned char*)d_packed, 1.0f);
  else
    fused_pack_kernel<tc::FMT_BF16, false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_bond_transform, *w, (unsigned char*)d_packed, 1.0f);
  IMP_LAUNCH_CHECK();
  return 0;
}

namespace imp {
int launch_fused_h4(const FusedArgs& a, int grid, bool precise, bool compact, cudaStream_t st);  // fused_fwd4.cu
}

static int fused_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

template <int FMT, bool PRECISE, int MP>
static int launch_fused(const FusedArgs& a, int grid, size_t smem, cudaStream_t stream) {
  IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_kernel<FMT, PRECISE, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mpnn_fused_kernel<FMT, PRECISE, MP><<<grid, 256, smem, stream>>>(a);
  IMP_LAUNCH_CHECK();
  return 0;
}

static int fused_forward_impl(const imp_graph_t* g, const imp_compact_graph_t* cg, const float* d_atom_emb, int32_t atom_vocab,
                              const float* d_bond_emb, int32_t d, int32_t bond_dim, int32_t steps, const void* d_packed,
                              float eps, int32_t flags, int32_t max_mol_atoms, float* d_pooled, int32_t* d_status, void* stream);

extern "C" int imp_mpnn_forward_fused(const imp_graph_t* g, const float* d_atom_emb, int32_t atom_vocab,
                                      const float* d_bond_emb, int32_t d, int32_t bond_dim, int32_t steps,
                                      const void* d_packed, float eps, int32_t flags, int32_t max_mol_atoms,
                                      float* d_pooled, int32_t* d_status, void* stream) {
  return fused_forward_impl(g, nullptr, d_atom_emb, atom_vocab, d_bond_emb, d, bond_dim, steps, d_packed, eps, flags,
                            max_mol_atoms, d_pooled, d_status, stream);
}

extern "C" int imp_mpnn_forward_fused_compact(const imp_compact_graph_t* cg, const float* d_atom_emb, int32_t atom_vocab,
                                              const float* d_bond_emb, int32_t d, int32_t bond_dim, int32_t steps,
                                              const void* d_packed, float eps, int32_t flags, int32_t max_mol_atoms,
                                              float* d_pooled, int32_t* d_status, void* stream) {
  IMP_REQUIRE(cg, IMP_ERR_ARG, "imp_mpnn_forward_fused_compact: graph is null");
  IMP_REQUIRE((flags & IMP_TC_FP16) && !(flags & (IMP_TC_F32_ZBUILD | IMP_TC_TWO_THREADS_PER_ROW)), IMP_ERR_UNSUPPORTED,
              "imp_mpnn_forward_fused_compact: the compact feed is read by the default half-operand kernel only");
  IMP_REQUIRE(atom_vocab <= 256 && cg->bond_vocab <= 256, IMP_ERR_DIM, "imp_mpnn_forward_fused_compact: vocabularies must fit 8 bits");
  imp_graph_t g{};
  g.n_pairs = cg->n_pairs, g.n_atoms = cg->n_atoms, g.n_cat_atoms = cg->n_cat_atoms, g.n_unique = cg->n_unique;
  g.n_edges = cg->n_edges, g.bond_vocab = cg->bond_vocab, g.mol_ptr = const_cast<int32_t*>(cg->mol_ptr);
  return fused_forward_impl(&g, cg, d_atom_emb, atom_vocab, d_bond_emb, d, bond_dim, steps, d_packed, eps, flags, max_mol_atoms,
                            d_pooled, d_status, stream);
}

static int fused_forward_impl(const imp_graph_t* g, const imp_compact_graph_t* cg, const float* d_atom_emb, int32_t atom_vocab,
                              const float* d_bond_emb, int32_t d, int32_t bond_dim, int32_t steps, const void* d_packed,
                              float eps, int32_t flags, int32_t max_mol_atoms, float* d_pooled, int32_t* d_status, void* stream) {
  const bool compact = cg != nullptr;
  IMP_REQUIRE(g, IMP_ERR_ARG, "imp_mpnn_forward_fused: graph is null");
  IMP_REQUIRE(g->n_pairs >= 0 && g->n_atoms >= 0 && g->n_cat_atoms >= 0 && g->n_cat_atoms <= g->n_atoms, IMP_ERR_ARG,
              "imp_mpnn_forward_fused: bad graph sizes");
  IMP_REQUIRE(d == FZ_D && bond_dim == FZ_K, IMP_ERR_DIM,
              "imp_mpnn_forward_fused: built for atom_dim %d, bond_dim %d (got %d, %d); use the staged kernels", FZ_D, FZ_K, d, bond_dim);
  IMP_REQUIRE(steps >= 1 && steps <= FZ_MAX_STEPS, IMP_ERR_DIM, "imp_mpnn_forward_fused: 1..%d steps (got %d)", FZ_MAX_STEPS, steps);
  IMP_REQUIRE(g->bond_vocab >= 1 && g->bond_vocab <= FZ_MAX_VB && atom_vocab >= 1, IMP_ERR_DIM,
              "imp_mpnn_forward_fused: bond vocabulary must be in 1..%d", FZ_MAX_VB);
  IMP_REQUIRE(max_mol_atoms <= FZ_ROWS, IMP_ERR_DIM,
              "imp_mpnn_forward_fused: a molecule has %d atoms, a tile holds %d; use the staged kernels", max_mol_atoms, FZ_ROWS);
  if (g->n_pairs == 0) return 0;
  IMP_REQUIRE(d_atom_emb && d_bond_emb && d_packed && d_pooled && g->mol_ptr, IMP_ERR_ARG, "imp_mpnn_forward_fused: null pointer");
  if (compact) {
    IMP_REQUIRE(cg->mol_eptr && cg->atom_w && (g->n_unique == 0 || cg->edge_w), IMP_ERR_ARG,
                "imp_mpnn_forward_fused_compact: null index arrays");
  } else {
    IMP_REQUIRE(g->atom_id && g->row_ptr, IMP_ERR_ARG, "imp_mpnn_forward_fused: null pointer");
    IMP_REQUIRE(g->n_unique == 0 || (g->col_src && g->edge_bm), IMP_ERR_ARG, "imp_mpnn_forward_fused: null edge arrays");
  }
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_mpnn_forward_fused: tcgen05 needs an sm_100 device");
  FusedArgs a;
  a.mol_ptr = g->mol_ptr, a.atom_id = g->atom_id, a.row_ptr = g->row_ptr, a.col_src = g->col_src, a.edge_bm = g->edge_bm;
  a.atom_emb = d_atom_emb, a.bond_emb = d_bond_emb, a.packed = (const unsigned char*)d_packed, a.pooled = d_pooled;
  a.n_atoms = g->n_atoms, a.n_unique = g->n_unique;

  a.status = d_status, a.n_pairs = g->n_pairs, a.atom_vocab = atom_vocab, a.bond_vocab = g->bond_vocab, a.steps = steps, a.eps = eps;
  a.prof = nullptr;
  a.mol_eptr = nullptr, a.atom_w = nullptr, a.edge_w = nullptr;
  if (compact) a.mol_eptr = cg->mol_eptr, a.atom_w = cg->atom_w, a.edge_w = cg->edge_w;
#ifdef FZ_PROFILE
  a.debug = (flags >> 16) & 0xff;  // timing ablations, profiling builds only
#else
  a.debug = 0;
  IMP_REQUIRE((flags & ~0x7ff) == 0, IMP_ERR_ARG, "imp_mpnn_forward_fused: unknown flag bits 0x%x", flags & ~0x7ff);
#endif
#ifdef FZ_PROFILE
  a.prof = reinterpret_cast<long long*>(d_status);  // profiling build: d_status must hold 3 * 32 int64 (zeroed by the caller)
  a.status = nullptr;
#endif
  // one persistent CTA per SM; CTAs are split between the towers in proportion to their atoms
  const int sms = fused_sm_count();
  const int n_groups = (int)ceil_div(g->n_pairs, FZ_GROUP);
  const int want = (int)ceil_div(n_groups, 2);  // two warpgroups per CTA
  int n_cat = (int)((int64_t)sms * g->n_cat_atoms / (g->n_atoms > 0 ? g->n_atoms : 1));
  n_cat = n_cat < 1 ? 1 : (n_cat > sms - 1 ? sms - 1 : n_cat);
  int n_an = sms - n_cat;
  if (n_cat > want) n_cat = want;
  if (n_an > want) n_an = want;
  a.n_cta_cat = n_cat;
  const int grid = n_cat + n_an;
  const size_t smem = (size_t)fused_smem_bytes(steps, g->bond_vocab);
  IMP_REQUIRE(smem <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused: needs %zu B of shared memory", smem);
  const bool f16 = flags & IMP_TC_FP16, precise = flags & IMP_TC_PRECISE_EPILOGUE, mp8 = flags & IMP_TC_MP8;
  cudaStream_t st = (cudaStream_t)stream;
  if (f16 && !(flags & IMP_TC_F32_ZBUILD)) {
    if (flags & IMP_TC_TWO_THREADS_PER_ROW) {  // second generation: 2 contexts x 256 threads
      const size_t smem2 = (size_t)fused2_smem_bytes(steps, g->bond_vocab);
      IMP_REQUIRE(smem2 <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused: needs %zu B of shared memory", smem2);
      // work units are distributed over 2 contexts per CTA
      if (precise) {
        IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        mpnn_fused_h2_kernel<true><<<grid, 2 * F2_CTX_THREADS, smem2, st>>>(a);
      } else {
        IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        mpnn_fused_h2_kernel<false><<<grid, 2 * F2_CTX_THREADS, smem2, st>>>(a);
      }
      IMP_LAUNCH_CHECK();
      return 0;
    }
    // default for half operands: third generation, 4 contexts x 128 threads (IMP_TC_THREE_CONTEXTS: 3 contexts);
    // IMP_TC_GEN4 selects the fourth generation (fused_fwd4.cu: no blocking barrier in a step) for comparison
    const int nctx = (flags & IMP_TC_THREE_CONTEXTS) ? 3 : 4;
    if ((flags & IMP_TC_GEN4) && !(flags & (IMP_TC_GEN3 | IMP_TC_THREE_CONTEXTS))) {
      int nc4 = (int)((int64_t)sms * g->n_cat_atoms / (g->n_atoms > 0 ? g->n_atoms : 1));
      nc4 = nc4 < 1 ? 1 : (nc4 > sms - 1 ? sms - 1 : nc4);
      int na4 = sms - nc4;
      const int want4 = (int)ceil_div(n_groups, 4);
      if (nc4 > want4) nc4 = want4;
      if (na4 > want4) na4 = want4;
      a.n_cta_cat = nc4;
      return launch_fused_h4(a, nc4 + na4, precise, compact, st);
    }
    int nc = (int)((int64_t)sms * g->n_cat_atoms / (g->n_atoms > 0 ? g->n_atoms : 1));
    nc = nc < 1 ? 1 : (nc > sms - 1 ? sms - 1 : nc);
    int na = sms - nc;
    const int want3 = (int)ceil_div(n_groups, nctx);
    if (nc > want3) nc = want3;
    if (na > want3) na = want3;
    a.n_cta_cat = nc;
    const int grid3 = nc + na;
    const size_t smem3 = (size_t)fused3_smem_bytes(steps, g->bond_vocab, nctx);
    IMP_REQUIRE(smem3 <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused: needs %zu B of shared memory", smem3);
#define IMP_LAUNCH_H2X(PREC, NC)                                                                                           \
  do {                                                                                                                     \
    if (compact) {                                                                                                         \
      IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h2x_kernel<PREC, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
      mpnn_fused_h2x_kernel<PREC, NC, true><<<grid3, NC * F3_CTX_THREADS, smem3, st>>>(a);                                \
    } else {                                                                                                               \
      IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h2x_kernel<PREC, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
      mpnn_fused_h2x_kernel<PREC, NC, false><<<grid3, NC * F3_CTX_THREADS, smem3, st>>>(a);                               \
    }                                                                                                                      \
  } while (0)
    if (nctx == 4) {
      if (precise) IMP_LAUNCH_H2X(true, 4); else IMP_LAUNCH_H2X(false, 4);
    } else {
      if (precise) IMP_LAUNCH_H2X(true, 3); else IMP_LAUNCH_H2X(false, 3);
    }
#undef IMP_LAUNCH_H2X
    IMP_LAUNCH_CHECK();
    return 0;
  }
  if (f16) {
    if (precise) return mp8 ? launch_fused<tc::FMT_F16, true, 8>(a, grid, smem, st) : launch_fused<tc::FMT_F16, true, 16>(a, grid, smem, st);
    return mp8 ? launch_fused<tc::FMT_F16, false, 8>(a, grid, smem, st) : launch_fused<tc::FMT_F16, false, 16>(a, grid, smem, st);
  }
  if (precise) return mp8 ? launch_fused<tc::FMT_BF16, true, 8>(a, grid, smem, st) : launch_fused<tc::FMT_BF16, true, 16>(a, grid, smem, st);
  return mp8 ? launch_fused<tc::FMT_BF16, false, 8>(a, grid, smem, st) : launch_fused<tc::FMT_BF16, false, 16>(a, grid, smem, st);
}
