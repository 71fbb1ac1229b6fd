// Fifth generation of the fused whole-tower forward ("h5", the default for IEEE-half operands): the third generation's
// step pipeline (fused_fwd.cu: Embedding -> [BondMatrixMessage o Reduce -> GatedUpdate] x S -> GlobalSumPool,
// train_viscosity.py:163-187, models/layers.py:57-164; four 128-thread contexts per CTA, one thread per atom row, packed
// HFMA2 Z build in two K halves into tensor memory, three tcgen05 GEMM groups per step, fp32 state in registers) fed from
// the tile plan of fused_plan.cuh instead of the CSR arrays:
//   * a tile's indices arrive as ONE 2 KiB TMA bulk copy, double-buffered (the next tile's record is in flight during the
//     current tile's steps); the in-kernel group scan, in-degree sort, entry decode and their five context barriers are gone;
//   * tiles are ~97 % full (best-fit over a 256-molecule window) instead of ~88 %;
//   * the Z accumulators start from the first entry's products (no zero fill + FMA).
#include "fused_common.cuh"
#include "fused_plan.cuh"

namespace imp {

constexpr int F5_CTX = 4;
constexpr int F5_THREADS = 128;

struct alignas(128) FusedWgSmem5 {
  uint32_t hb[FZ_ROWS * FZ_HS];  // words 0..15 of a row: h as packed halves; after the last step: fp32 h rows for the pooling
  FusedTile plan[2];
  uint64_t bar[4];   // 0: GEMM1 done, 1: GEMM2 done, 2: GEMM3 done, 3: GEMM1a done
  uint64_t pbar[2];  // plan buffers
  uint64_t pad[10];
};

__host__ __device__ inline int fused5_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack::BYTES + ctab + F5_CTX * (int)sizeof(FusedWgSmem5) + (int)sizeof(FusedCtl);
}

struct Fused5Args {
  const unsigned char* plan;
  const float* atom_emb;
  const float* bond_emb;
  const unsigned char* packed;  // [2][steps][FusedPack::BYTES]
  float* pooled;                // [2P][32]
  int atom_vocab, bond_vocab, steps, n_cta_cat;
  float eps;
};

struct F5True { static constexpr bool value = true; };
struct F5False { static constexpr bool value = false; };

template <bool PRECISE>
__global__ void __launch_bounds__(F5_CTX * F5_THREADS, 1) mpnn_fused_h5_kernel(const Fused5Args a) {
  constexpr int D = FZ_D;
  constexpr int FMT = tc::FMT_F16;
  constexpr int NT = F5_CTX * F5_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 7, t = tid & 127, wq = warp & 3;
  const int wbytes = a.steps * FusedPack::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);
  FusedWgSmem5& ws = reinterpret_cast<FusedWgSmem5*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + F5_CTX * sizeof(FusedWgSmem5));

  const int tower = blockIdx.x >= a.n_cta_cat;
  const FusedPlanHeader* hdr = reinterpret_cast<const FusedPlanHeader*>(a.plan);
  if (__ldg(&hdr->status) == 2) return;  // the plan ran out of tile records: some are unwritten (the host raises, model.check_status)
  const int n_tiles = min(__ldg(&hdr->n_tiles[tower]), __ldg(&hdr->cap[tower]));
  const FusedTile* tiles = reinterpret_cast<const FusedTile*>(a.plan + FP_HEADER_BYTES) + (size_t)(tower ? __ldg(&hdr->cap[0]) : 0);
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const int first = cta_in_tower * F5_CTX + ctx, stride = n_cta_tower * F5_CTX;

  if (tid == 0) {  // resident weights of this tower (all steps): one TMA bulk copy
    tc::mbar_init(&ctl.wbar, 1);
    tc::mbar_fence_init();
    tc::mbar_arrive_expect_tx(&ctl.wbar, (uint32_t)wbytes);
    tc::bulk_copy_g2s(smem, a.packed + (size_t)tower * wbytes, (uint32_t)wbytes, &ctl.wbar);
  }
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::mbar_init(&ws.bar[i], 1);
    tc::mbar_init(&ws.pbar[0], 1);
    tc::mbar_init(&ws.pbar[1], 1);
    tc::mbar_fence_init();
    if (first < n_tiles) {  // first tile record of this context
      tc::mbar_arrive_expect_tx(&ws.pbar[0], (uint32_t)sizeof(FusedTile));
      tc::bulk_copy_g2s(&ws.plan[0], tiles + first, (uint32_t)sizeof(FusedTile), &ws.pbar[0]);
    }
  }
  for (int i = tid; i < a.bond_vocab; i += NT) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i + 1);
    s_ctab[i] = make_uint4(tc::pack_f16x2(c0.x, c0.y), tc::pack_f16x2(c0.z, c0.w), tc::pack_f16x2(c1.x, c1.y),
                           tc::pack_f16x2(c1.z, c1.w));
  }
  if (warp == 0) tc::tmem_alloc<512>(&ctl.tmem_base);
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  tc::mbar_wait(&ctl.wbar, 0);

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
  const bool mma_warp = wq == 0;
  const int bar_id = 1 + ctx;
  const int myslot = (ctx & 1) ? FZ_ROWS - 1 - t : t;  // alternate contexts walk the in-degree order in opposite directions
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  uint32_t ph = 0, pph = 0;  // parities: per-step MMA barriers; plan buffers (bit b = buffer b)

  int buf = 0;
  for (int tile = first; tile < n_tiles; tile += stride, buf ^= 1) {
    if (t == 0 && tile + stride < n_tiles) {  // next record -> the other buffer (its readers passed the end-of-tile barrier)
      tc::fence_proxy_async_smem();
      tc::mbar_arrive_expect_tx(&ws.pbar[buf ^ 1], (uint32_t)sizeof(FusedTile));
      tc::bulk_copy_g2s(&ws.plan[buf ^ 1], tiles + tile + stride, (uint32_t)sizeof(FusedTile), &ws.pbar[buf ^ 1]);
    }
    tc::mbar_wait(&ws.pbar[buf], (pph >> buf) & 1u);
    pph ^= 1u << buf;
    const FusedTile& tp = ws.plan[buf];
    const uint32_t sw = tp.slot[myslot];
    const int r = sw & 127, deg = (sw >> 7) & 31, aid = (int)(sw >> 22);
    const uint32_t* entp = &tp.ent[(sw >> 12) & 1023];
    uint32_t* hbrow = &ws.hb[r * FZ_HS];
    float h[D];
    {  // Embedding(atom)
      const float4* er = emb4 + aid * (D / 4);
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 x = __ldg(er + c);
        h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
        reinterpret_cast<uint2*>(hbrow)[c] = make_uint2(tc::pack_f16x2(x.x, x.y), tc::pack_f16x2(x.z, x.w));
      }
    }
    tc::named_bar_sync(bar_id, F5_THREADS);

    for (int s = 0; s < a.steps; ++s) {
      const uint64_t dstep = (uint64_t)(s * (FusedPack::BYTES / 16));
      const float* bias = reinterpret_cast<const float*>(smem + s * FusedPack::BYTES + FusedPack::OFF_BIAS);
      // ------------------------------------------------------------ Z in two K halves -> TMEM -> GEMM1
#pragma unroll 1
      for (int hz = 0; hz < 2; ++hz) {
        __half2 acc[D * 2];
        // one entry: acc (+)= h[src] (x) (mult * c[4 hz .. 4 hz + 4))
        auto entry = [&](uint32_t ec, auto first_entry) {
          const uint2 cu = reinterpret_cast<const uint2*>(s_ctab + ((ec >> 8) & 0xff))[hz];
          const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
          const __half2 mult = *reinterpret_cast<const __half2*>(&mbits);
          const __half2 c0 = __hmul2(*reinterpret_cast<const __half2*>(&cu.x), mult);
          const __half2 c1 = __hmul2(*reinterpret_cast<const __half2*>(&cu.y), mult);
          const uint4* hp = reinterpret_cast<const uint4*>(&ws.hb[(ec & 0x7f) * FZ_HS]);
#pragma unroll
          for (int q = 0; q < D / 8; ++q) {  // 8 columns per 16-byte read; HFMA2 broadcasts the low / high half
            const uint4 hv = hp[q];
            const __half2 hw[4] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y),
                                   *reinterpret_cast<const __half2*>(&hv.z), *reinterpret_cast<const __half2*>(&hv.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __half2 lo = __low2half2(hw[i]), hi = __high2half2(hw[i]);
              const int m = 8 * q + 2 * i;
              if constexpr (decltype(first_entry)::value) {
                acc[m * 2] = __hmul2(lo, c0), acc[m * 2 + 1] = __hmul2(lo, c1);
                acc[m * 2 + 2] = __hmul2(hi, c0), acc[m * 2 + 3] = __hmul2(hi, c1);
              } else {
                acc[m * 2] = __hfma2(lo, c0, acc[m * 2]), acc[m * 2 + 1] = __hfma2(lo, c1, acc[m * 2 + 1]);
                acc[m * 2 + 2] = __hfma2(hi, c0, acc[m * 2 + 2]), acc[m * 2 + 3] = __hfma2(hi, c1, acc[m * 2 + 3]);
              }
            }
          }
        };
        {
          // a row without entries runs the first-entry code on the all-zero descriptor (multiplicity 0: every product is 0):
          // a branch here is if-converted into 64 selects per half for EVERY warp (9 % of the kernel's instructions, ncu)
          const uint32_t e_first = deg > 0 ? entp[0] : 0u;
          uint32_t en = deg > 1 ? entp[1] : 0u;
          entry(e_first, F5True{});
#pragma unroll 1
          for (int e = 1; e < deg; ++e) {
            const uint32_t ec = en;
            if (e + 1 < deg) en = entp[e + 1];  // next entry's descriptor is in flight during this one's FMAs
            entry(ec, F5False{});
          }
        }
        if (hz == 1) {  // GEMM1a must have consumed the first half before its columns are rewritten
          tc::mbar_wait(&ws.bar[3], ph);
          tc::fence_after_thread_sync();
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = *reinterpret_cast<const uint32_t*>(&acc[ch * 32 + i]);
          tc::tmem_st32(tZ + lane_off + (uint32_t)(ch * 32), rr);
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        tc::named_bar_sync(bar_id, F5_THREADS);
        if (mma_warp) {
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
      tc::mbar_wait(&ws.bar[0], ph);
      tc::fence_after_thread_sync();
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
      tc::fence_before_thread_sync();
      tc::named_bar_sync(bar_id, F5_THREADS);
      // ------------------------------------------------------------ GEMM2: 0.5 ([h | agg | 1] . [Wz | Wr ; bz | br])
      if (mma_warp) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma_f16_ts(tCzr, tAh + 8 * ks, dBzr + dstep + (uint64_t)(ks * 128), idesc64, ks > 0);
          tc::mma_f16_ts(tCzr, tOnes, dBBzr + dstep, idesc64, true);
          tc::mma_commit(&ws.bar[1]);
        }
        __syncwarp();
      }
      tc::mbar_wait(&ws.bar[1], ph);
      tc::fence_after_thread_sync();
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
      tc::fence_before_thread_sync();
      tc::named_bar_sync(bar_id, F5_THREADS);
      // ------------------------------------------------------------ GEMM3: [agg | r*h | 1] . [Wh[d:2d] ; Wh[0:d] ; bh]
      if (mma_warp) {
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
      tc::mbar_wait(&ws.bar[2], ph);
      tc::fence_after_thread_sync();
      {  // candidate, blend, LayerNorm (biased variance, eps), residual  (models/layers.py:151-156)
        float gq[32];
        tc::tmem_ld32(tCht + lane_off, gq);
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < D; j += 2) {
          const float n0 = fmaf(z[j], fz_tanh<PRECISE>(gq[j]) - h[j], h[j]);
          const float n1 = fmaf(z[j + 1], fz_tanh<PRECISE>(gq[j + 1]) - h[j + 1], h[j + 1]);
          gq[j] = n0, gq[j + 1] = n1;
          s0 += n0, s1 += n1;
          q0 = fmaf(n0, n0, q0), q1 = fmaf(n1, n1, q1);
        }
        const float mean = (s0 + s1) * (1.0f / D);
        const float var = fmaxf(fmaf(q0 + q1, 1.0f / D, -mean * mean), 0.f);  // biased variance
        const float inv = PRECISE ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
        const float ninv = -mean * inv;
#pragma unroll
        for (int j = 0; j < D; ++j) h[j] = fmaf(fmaf(gq[j], inv, ninv), bias[3 * D + j], h[j]) + bias[4 * D + j];
        if (s + 1 < a.steps) {
#pragma unroll
          for (int c = 0; c < D / 8; ++c)
            reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(tc::pack_f16x2(h[8 * c], h[8 * c + 1]), tc::pack_f16x2(h[8 * c + 2], h[8 * c + 3]),
                                                            tc::pack_f16x2(h[8 * c + 4], h[8 * c + 5]), tc::pack_f16x2(h[8 * c + 6], h[8 * c + 7]));
        } else {  // fp32 rows for the pooling; id 0 is not pooled (models/layers.py:163)
          const float keep = aid > 0 ? 1.f : 0.f;
#pragma unroll
          for (int c = 0; c < D / 4; ++c)
            reinterpret_cast<float4*>(hbrow)[c] = make_float4(keep * h[4 * c], keep * h[4 * c + 1], keep * h[4 * c + 2], keep * h[4 * c + 3]);
        }
      }
      tc::fence_before_thread_sync();
      tc::named_bar_sync(bar_id, F5_THREADS);
      ph ^= 1;
    }
    // ---------------------------------------------------------------- GlobalSumPool
    {
      const float* hfp = reinterpret_cast<const float*>(ws.hb);
      const int nm = tp.nm;
      for (int mi = (t >> 5); mi < nm; mi += 4) {
        const int lo = tp.mol_lo[mi], hi = tp.mol_lo[mi + 1];
        // four interleaved partial sums, combined in a fixed order
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int rr = lo;
        for (; rr + 4 <= hi; rr += 4) {
          s0 += hfp[rr * FZ_HS + lane], s1 += hfp[(rr + 1) * FZ_HS + lane];
          s2 += hfp[(rr + 2) * FZ_HS + lane], s3 += hfp[(rr + 3) * FZ_HS + lane];
        }
        for (; rr < hi; ++rr) s0 += hfp[rr * FZ_HS + lane];
        a.pooled[(size_t)tp.molid[mi] * D + lane] = (s0 + s1) + (s2 + s3);
      }
    }
    tc::named_bar_sync(bar_id, F5_THREADS);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

}  // namespace imp

using namespace imp;

static int fused5_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

namespace imp {
int launch_fused_h6(const void* d_plan, int32_t n_atoms, int32_t n_cat_atoms, int32_t bond_vocab, const float* d_atom_emb,
                    int32_t atom_vocab, const float* d_bond_emb, int32_t steps, const void* d_packed, float eps, bool precise,
                    float* d_pooled, cudaStream_t st);  // fused_fwd6.cu
int launch_fused_h7(const void* d_plan, int32_t n_atoms, int32_t n_cat_atoms, int32_t bond_vocab, const float* d_atom_emb,
                    int32_t atom_vocab, const float* d_bond_emb, int32_t steps, const void* d_packed, float eps, bool precise,
                    float* d_pooled, cudaStream_t st);  // fused_fwd7.cu
int launch_fused_h8(const void* d_plan, int32_t n_atoms, int32_t n_cat_atoms, int32_t bond_vocab, const float* d_atom_emb,
                    int32_t atom_vocab, const float* d_bond_emb, int32_t steps, const void* d_packed, float eps, bool precise,
                    float* d_pooled, cudaStream_t st);  // fused_fwd8.cu
}

extern "C" int imp_mpnn_forward_fused_planned(const void* d_plan, int32_t n_pairs, int32_t n_atoms, int32_t n_cat_atoms,
                                              int32_t bond_vocab, const float* d_atom_emb, int32_t atom_vocab,
                                              const float* d_bond_emb, int32_t d, int32_t bond_dim, int32_t steps,
                                              const void* d_packed, float eps, int32_t flags, float* d_pooled, void* stream) {
  IMP_REQUIRE(n_pairs >= 0 && n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG,
              "imp_mpnn_forward_fused_planned: bad graph sizes");
  IMP_REQUIRE(d == FZ_D && bond_dim == FZ_K, IMP_ERR_DIM,
              "imp_mpnn_forward_fused_planned: built for atom_dim %d, bond_dim %d (got %d, %d); use the staged kernels", FZ_D, FZ_K, d,
              bond_dim);
  IMP_REQUIRE(steps >= 1 && steps <= FZ_MAX_STEPS, IMP_ERR_DIM, "imp_mpnn_forward_fused_planned: 1..%d steps (got %d)", FZ_MAX_STEPS, steps);
  IMP_REQUIRE(bond_vocab >= 1 && bond_vocab <= FZ_MAX_VB && atom_vocab >= 1 && atom_vocab <= 1024, IMP_ERR_DIM,
              "imp_mpnn_forward_fused_planned: bond vocabulary must be in 1..%d, atom vocabulary in 1..1024", FZ_MAX_VB);
  const int gen_flags = flags & (IMP_TC_GEN5 | IMP_TC_GEN7 | IMP_TC_GEN8);
  IMP_REQUIRE((flags & IMP_TC_FP16) && !(flags & ~(IMP_TC_FP16 | IMP_TC_PRECISE_EPILOGUE | IMP_TC_GEN5 | IMP_TC_GEN7 | IMP_TC_GEN8)) &&
                  (gen_flags & (gen_flags - 1)) == 0,
              IMP_ERR_UNSUPPORTED,
              "imp_mpnn_forward_fused_planned: IEEE-half operands only (flags IMP_TC_FP16 [| IMP_TC_PRECISE_EPILOGUE] [| one of "
              "IMP_TC_GEN5 / GEN7 / GEN8])");
  if (n_pairs == 0) return 0;
  IMP_REQUIRE(d_plan && d_atom_emb && d_bond_emb && d_packed && d_pooled, IMP_ERR_ARG, "imp_mpnn_forward_fused_planned: null pointer");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_mpnn_forward_fused_planned: tcgen05 needs an sm_100 device");
  if (flags & IMP_TC_GEN8)  // eighth generation (weights from imp_fused_pack_planned7)
    return launch_fused_h8(d_plan, n_atoms, n_cat_atoms, bond_vocab, d_atom_emb, atom_vocab, d_bond_emb, steps, d_packed, eps,
                           (flags & IMP_TC_PRECISE_EPILOGUE) != 0, d_pooled, (cudaStream_t)stream);
  if (flags & IMP_TC_GEN7)  // seventh generation (weights from imp_fused_pack_planned7)
    return launch_fused_h7(d_plan, n_atoms, n_cat_atoms, bond_vocab, d_atom_emb, atom_vocab, d_bond_emb, steps, d_packed, eps,
                           (flags & IMP_TC_PRECISE_EPILOGUE) != 0, d_pooled, (cudaStream_t)stream);
  if (!(flags & IMP_TC_GEN5))  // sixth generation (weights from imp_fused_pack_planned)
    return launch_fused_h6(d_plan, n_atoms, n_cat_atoms, bond_vocab, d_atom_emb, atom_vocab, d_bond_emb, steps, d_packed, eps,
                           (flags & IMP_TC_PRECISE_EPILOGUE) != 0, d_pooled, (cudaStream_t)stream);
  Fused5Args a;
  a.plan = (const unsigned char*)d_plan, a.atom_emb = d_atom_emb, a.bond_emb = d_bond_emb, a.packed = (const unsigned char*)d_packed;
  a.pooled = d_pooled, a.atom_vocab = atom_vocab, a.bond_vocab = bond_vocab, a.steps = steps, a.eps = eps;
  // one persistent CTA per SM; CTAs are split between the towers in proportion to their atoms
  const int sms = fused5_sm_count();
  int nc = (int)((int64_t)sms * n_cat_atoms / (n_atoms > 0 ? n_atoms : 1));
  nc = nc < 1 ? 1 : (nc > sms - 1 ? sms - 1 : nc);
  int na = sms - nc;
  const int want = (int)ceil_div(ceil_div((int64_t)n_atoms, 100), F5_CTX) + 1;  // never more CTAs than a small batch has tiles for
  if (nc > want) nc = want;
  if (na > want) na = want;
  a.n_cta_cat = nc;
  const size_t smem = (size_t)fused5_smem_bytes(steps, bond_vocab);
  IMP_REQUIRE(smem <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused_planned: needs %zu B of shared memory", smem);
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & IMP_TC_PRECISE_EPILOGUE) {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h5_kernel<true><<<nc + na, F5_CTX * F5_THREADS, smem, st>>>(a);
  } else {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h5_kernel<false><<<nc + na, F5_CTX * F5_THREADS, smem, st>>>(a);
  }
  IMP_LAUNCH_CHECK();
  return 0;
}
