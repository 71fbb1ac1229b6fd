// Eighth generation of the fused whole-tower forward ("h8"): generation 6 (fused_fwd6.cu: four 128-thread contexts, one thread
// per atom row in the gate phases, gate GEMMs on the in-place tf32 accumulator) with the Z rows built FOUR LANES WIDE as in
// generation 7 (fused_fwd7.cu).
//
// Replaces the same reference code: Embedding -> [BondMatrixMessage o Reduce -> GatedUpdate] x S -> GlobalSumPool
// (train_viscosity.py:163-187, models/layers.py:57-164).
//
// Phase timing of generation 6 (tools/fused_prof2.py): 46 % of a context's time is the Z phase -- 30 % accumulating, 16 %
// waiting at the operand barrier for the warp that owns the highest in-degrees -- and the shared-memory pipe is what the Z
// phases of the four contexts compete for (16-byte reads of 32 random rows conflict 2.15x; ncu: shared wavefronts 53 % of all
// cycles).  Here lane 4 g + j of warp q accumulates, for the rows in TMEM lanes g, g + 8 (then g + 16, g + 24) of quadrant q,
// the state columns 16 hz + 4 j .. + 3 of K half hz against the eight bond components: the four lanes of a row read 32
// contiguous bytes of the neighbour's row, the warp writes two rows' fragments with one tcgen05.st.16x256b, the in-degree
// granularity is 8 rows, and the plan's in-degree order is dealt out so that warp q gets octets q, 15 - q, 7 - q, 8 + q:
// the four warps of a tile carry the same number of entries.  The entries' coefficient vectors mult * c[bond] and source-row
// addresses are formed once per tile.  Weights: imp_fused_pack_planned7 (K order of the 16x256b fragments).
#include "fused_common.cuh"
#include "fused_pack6.cuh"
#include "fused_plan.cuh"
#include "fused_prof.cuh"

namespace imp {

constexpr int F8_CTX = 4;
constexpr int F8_THREADS = 128;
constexpr int F8_HS = 16;  // words per row of the shared-memory h copy (64 B)
#ifndef F8_ARRIVE
#define F8_ARRIVE 0  // 1: operand-ready barriers are arrivals on an mbarrier that only the MMA-issuing warp waits for
#endif

struct F8True { static constexpr bool value = true; };
struct F8False { static constexpr bool value = false; };

struct alignas(128) FusedWgSmem8 {
  uint32_t hb[FZ_ROWS * F8_HS];  // h as packed halves (natural row order); during the pooling: 16 fp32 columns of h
  uint4 cf[FP_ECAP + 1];         // per entry of the tile: mult * c[bond][0..8) as halves; cf[FP_ECAP] = 0 pads short rows
  FusedTile plan[2];             // (ent[] of the current record is rewritten in place: shared-memory address of the source row)
  uint64_t bar[4];   // 1: gate GEMM done, 2: candidate GEMM done, 3: GEMM1a done
  uint64_t obar[3];  // F8_ARRIVE: operands in tensor memory (one arrival per warp): Z half 0, Z half 1, r * h
  uint64_t pbar[2];  // plan buffers
  uint32_t zero_ha;  // address word of the padding item (row 0 of the h copy)
  uint32_t pad32;
  uint64_t pad[6];
};

__host__ __device__ inline int fused8_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack6::BYTES + ctab + F8_CTX * (int)sizeof(FusedWgSmem8) + (int)sizeof(FusedCtl);
}

struct Fused8Args {
  const unsigned char* plan;
  const float* atom_emb;
  const float* bond_emb;
  const unsigned char* packed;  // [2][steps][FusedPack6::BYTES]
  float* pooled;                // [2P][32]
  int atom_vocab, bond_vocab, steps, n_cta_cat;
  float eps;
};

template <bool PRECISE>
__global__ void __launch_bounds__(F8_CTX * F8_THREADS, 1) mpnn_fused_h8_kernel(const Fused8Args a) {
  constexpr int D = FZ_D;
  constexpr int NT = F8_CTX * F8_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 7, t = tid & 127, wq = warp & 3;
  F6_PROF_DECL;
  const int wbytes = a.steps * FusedPack6::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);
  FusedWgSmem8& ws = reinterpret_cast<FusedWgSmem8*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + F8_CTX * sizeof(FusedWgSmem8));

  const int tower = blockIdx.x >= a.n_cta_cat;
  const FusedPlanHeader* hdr = reinterpret_cast<const FusedPlanHeader*>(a.plan);
  if (__ldg(&hdr->status) == 2) return;  // the plan ran out of tile records: some are unwritten (the host raises, model.check_status)
  const int n_tiles = min(__ldg(&hdr->n_tiles[tower]), __ldg(&hdr->cap[tower]));
  const FusedTile* tiles = reinterpret_cast<const FusedTile*>(a.plan + FP_HEADER_BYTES) + (size_t)(tower ? __ldg(&hdr->cap[0]) : 0);
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const int first = cta_in_tower * F8_CTX + ctx, stride = n_cta_tower * F8_CTX;

  if (tid == 0) {  // resident weights of this tower (all steps): one TMA bulk copy
    tc::mbar_init(&ctl.wbar, 1);
    tc::mbar_fence_init();
    tc::mbar_arrive_expect_tx(&ctl.wbar, (uint32_t)wbytes);
    tc::bulk_copy_g2s(smem, a.packed + (size_t)tower * wbytes, (uint32_t)wbytes, &ctl.wbar);
  }
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::mbar_init(&ws.bar[i], 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) tc::mbar_init(&ws.obar[i], F8_THREADS / 32);
    ws.cf[FP_ECAP] = make_uint4(0u, 0u, 0u, 0u);
    ws.zero_ha = tc::smem_u32(ws.hb);
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
  const uint32_t tZ = tbase, tCzr = tbase, tCht = tbase, tCagg = tbase + 64, tAh = tbase + 96, tOnes = tbase + 112;
  const uint32_t id32h = tc::make_idesc(tc::FMT_F16, FZ_ROWS, D), id64h = tc::make_idesc(tc::FMT_F16, FZ_ROWS, 2 * D);
  const uint32_t id32t = tc::make_idesc(tc::FMT_TF32, FZ_ROWS, D), id64t = tc::make_idesc(tc::FMT_TF32, FZ_ROWS, 2 * D);
  const uint64_t dWc = tc::make_smem_desc(sw0, D * 16, 128);
  const uint64_t dBzrh = tc::make_smem_desc(sw0 + FusedPack6::OFF_BZRH, 2 * D * 16, 128);
  const uint64_t dBzra = tc::make_smem_desc(sw0 + FusedPack6::OFF_BZRA, 2 * D * 16, 128);
  const uint64_t dBhh = tc::make_smem_desc(sw0 + FusedPack6::OFF_BHH, D * 16, 128);
  const uint64_t dBha = tc::make_smem_desc(sw0 + FusedPack6::OFF_BHA, D * 16, 128);
  const uint64_t dBBzr = tc::make_smem_desc(sw0 + FusedPack6::OFF_BBZR, 2 * D * 16, 128);
  const uint64_t dBBh = tc::make_smem_desc(sw0 + FusedPack6::OFF_BBH, D * 16, 128);
  const bool mma_warp = wq == ctx;  // one MMA-issuing warp per sub-partition (warp wq of every context runs on sub-partition wq)
  const int bar_id = 1 + ctx, opbar_id = 5 + ctx;
  // "my operand rows are in tensor memory"
  auto operands_ready = [&](int which) {
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
#if F8_ARRIVE
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&ws.obar[which]);
#else
    tc::named_bar_sync(opbar_id, F8_THREADS);
#endif
  };
  // Row of this thread in the gate phases: TMEM lane 32 wq + lane.  The plan lists the rows by in-degree (ascending); warp q
  // takes octets o, 15 - o (lanes 0-15: first 16x256b store) and 7 - o, 8 + o (lanes 16-31) of that order, o = (q + ctx) % 4.
  const int g = lane >> 2, j = lane & 3;
  const int oq = (wq + ctx) & 3;  // rotated by context: the four warps of a sub-partition carry the four different octet sets
  const int oct0 = oq, oct1 = 15 - oq, oct2 = 7 - oq, oct3 = 8 + oq;
  const int myoct = (lane & 16) ? ((lane & 8) ? oct3 : oct2) : ((lane & 8) ? oct1 : oct0);
  const int myslot = 8 * myoct + (lane & 7);
  const uint32_t hb_s = tc::smem_u32(ws.hb), cf_s = tc::smem_u32(ws.cf), zcf_s = cf_s + 16u * FP_ECAP, zha_s = tc::smem_u32(&ws.zero_ha);
  const uint32_t j8 = 8u * (uint32_t)j;
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  uint32_t ph = 0, pph = 0;  // parities: per-step MMA barriers; plan buffers (bit b = buffer b)
  {  // the constant (1, 0, ..., 0) K-step that carries the biases: written once, never overwritten
    const uint32_t ones[8] = {0x00003c00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tc::tmem_st8(tOnes + lane_off, ones);
    tc::tmem_wait_st();
  }

  int buf = 0;
  for (int tile = first; tile < n_tiles; tile += stride, buf ^= 1) {
    if (t == 0 && tile + stride < n_tiles) {  // next record -> the other buffer (its readers passed the end-of-tile barrier)
      tc::fence_proxy_async_smem();
      tc::mbar_arrive_expect_tx(&ws.pbar[buf ^ 1], (uint32_t)sizeof(FusedTile));
      tc::bulk_copy_g2s(&ws.plan[buf ^ 1], tiles + tile + stride, (uint32_t)sizeof(FusedTile), &ws.pbar[buf ^ 1]);
    }
    F6_PROF(0);
    tc::mbar_wait(&ws.pbar[buf], (pph >> buf) & 1u);
    pph ^= 1u << buf;
    F6_PROF(1);
    FusedTile& tp = ws.plan[buf];
    const uint32_t sw = tp.slot[myslot];
    const int r = sw & 127, aid = (int)(sw >> 22);
    // the four rows this thread accumulates in the Z phases (one of each of the warp's octets): in-degree | first entry << 8
    const uint32_t ent_s = tc::smem_u32(tp.ent);
    uint32_t zr[4];
    int dmax[4];
    {
      const int octs[4] = {oct0, oct1, oct2, oct3};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w = tp.slot[8 * octs[i] + g];
        zr[i] = ((w >> 7) & 31u) | (((w >> 12) & 1023u) << 8);
        dmax[i] = __reduce_max_sync(0xffffffffu, (int)((w >> 7) & 31u));  // octet maximum: warp-uniform trip count
      }
    }
    uint32_t* hbrow = &ws.hb[r * F8_HS];
    float h[D];
    {  // Embedding(atom): fp32 state in registers; packed once for the shared-memory copy (gathers) and the GEMM operand
      const float4* er = emb4 + aid * (D / 4);
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 x = __ldg(er + c);
        h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
        pk[2 * c] = tc::pack_f16x2(x.x, x.y), pk[2 * c + 1] = tc::pack_f16x2(x.z, x.w);
      }
#pragma unroll
      for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      tc::tmem_st16(tAh + lane_off, pk);
      tc::tmem_wait_st();
    }
    for (int e = t; e < (int)tp.n_ent; e += F8_THREADS) {  // once per tile: coefficient vector and source-row address of every entry
      const uint32_t ec = tp.ent[e];
      const uint4 c = s_ctab[(ec >> 8) & 0xff];
      const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
      const __half2 mult = *reinterpret_cast<const __half2*>(&mbits);
      __half2 p0 = __hmul2(*reinterpret_cast<const __half2*>(&c.x), mult), p1 = __hmul2(*reinterpret_cast<const __half2*>(&c.y), mult);
      __half2 p2 = __hmul2(*reinterpret_cast<const __half2*>(&c.z), mult), p3 = __hmul2(*reinterpret_cast<const __half2*>(&c.w), mult);
      ws.cf[e] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                            *reinterpret_cast<uint32_t*>(&p3));
      tp.ent[e] = hb_s + (ec & 0x7fu) * (F8_HS * 4);
    }
    tc::fence_before_thread_sync();
    F6_PROF(2);
    tc::named_bar_sync(bar_id, F8_THREADS);
    F6_PROF(3);

    for (int s = 0; s < a.steps; ++s) {
      const uint64_t dstep = (uint64_t)(s * (FusedPack6::BYTES / 16));
      const float* gb = reinterpret_cast<const float*>(smem + s * FusedPack6::BYTES + FusedPack6::OFF_BIAS);
      // ------------------------------------------------------------ Z in two K halves -> TMEM -> GEMM1 (-> gate GEMM)
#pragma unroll 1
      for (int hz = 0; hz < 2; ++hz) {
        const uint32_t hoff = 32u * (uint32_t)hz + j8;
        // one item of one row: acc (+)= h[src][16 hz + 4 j .. + 4) (x) (mult * c[0..8)); acc[4 c + i]: state column 16 hz + 4 j + c,
        // bond components 2 i, 2 i + 1
        auto fma_item = [&](const uint4 cf, const uint2 hv, __half2(&acc)[16], auto first_entry) {
          const __half2 c[4] = {*reinterpret_cast<const __half2*>(&cf.x), *reinterpret_cast<const __half2*>(&cf.y),
                                *reinterpret_cast<const __half2*>(&cf.z), *reinterpret_cast<const __half2*>(&cf.w)};
          const __half2 hw[2] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y)};
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const __half2 lo = __low2half2(hw[i]), hi = __high2half2(hw[i]);
#pragma unroll
            for (int kp = 0; kp < 4; ++kp) {
              if constexpr (decltype(first_entry)::value) {
                acc[8 * i + kp] = __hmul2(lo, c[kp]), acc[8 * i + 4 + kp] = __hmul2(hi, c[kp]);
              } else {
                acc[8 * i + kp] = __hfma2(lo, c[kp], acc[8 * i + kp]), acc[8 * i + 4 + kp] = __hfma2(hi, c[kp], acc[8 * i + 4 + kp]);
              }
            }
          }
        };
        // one row: the octet's trip count is warp-uniform; a row with fewer entries runs the rest on the zero item.
        // (Two rows side by side in one loop -- two independent load -> FMA chains -- measured 2 % slower: 9.04 vs 8.83 ms.)
        auto zrow = [&](const uint32_t rowinfo, const int dm, __half2(&acc)[16]) {
          const int deg = (int)(rowinfo & 0xffu);
          uint32_t pc = deg > 0 ? cf_s + 16u * (rowinfo >> 8) : zcf_s, pa = deg > 0 ? ent_s + 4u * (rowinfo >> 8) : zha_s;
          uint4 cf = tc::lds128(pc);
          uint32_t ha = tc::lds32(pa);
          {
            const uint2 hv = tc::lds64(ha + hoff);
            const bool more = 1 < deg;
            pc = more ? pc + 16u : zcf_s, pa = more ? pa + 4u : zha_s;
            const uint4 ncf = tc::lds128(pc);  // the next item is in flight during this one's FMAs
            ha = tc::lds32(pa);
            fma_item(cf, hv, acc, F8True{});
            cf = ncf;
          }
#pragma unroll 1
          for (int e = 1; e < dm; ++e) {
            const uint2 hv = tc::lds64(ha + hoff);
            const bool more = e + 1 < deg;
            pc = more ? pc + 16u : zcf_s, pa = more ? pa + 4u : zha_s;
            const uint4 ncf = tc::lds128(pc);
            ha = tc::lds32(pa);
            fma_item(cf, hv, acc, F8False{});
            cf = ncf;
          }
        };
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {  // lanes 0-15, then 16-31 of the quadrant
          __half2 accA[16], accB[16];
          zrow(zr[2 * pass], dmax[2 * pass], accA);
          zrow(zr[2 * pass + 1], dmax[2 * pass + 1], accB);
          F6_PROF(4);
          if (hz == 1 && pass == 0) {  // GEMM1a must have consumed the first half before its columns are rewritten
            tc::mbar_wait(&ws.bar[3], ph);
            tc::fence_after_thread_sync();
          }
          F6_PROF(5);
          uint32_t rr[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {  // column group c = 2 (state column) + (bond component / 4): accumulators 2 c, 2 c + 1
            rr[4 * c] = *reinterpret_cast<const uint32_t*>(&accA[2 * c]), rr[4 * c + 1] = *reinterpret_cast<const uint32_t*>(&accA[2 * c + 1]);
            rr[4 * c + 2] = *reinterpret_cast<const uint32_t*>(&accB[2 * c]), rr[4 * c + 3] = *reinterpret_cast<const uint32_t*>(&accB[2 * c + 1]);
          }
          tc::tmem_st_16x256b_x8(tZ + lane_off + ((uint32_t)(16 * pass) << 16), rr);
        }
        F6_PROF(6);
        operands_ready(hz);
        F6_PROF(7);
        if (mma_warp) {
#if F8_ARRIVE
          tc::mbar_wait(&ws.obar[hz], ph);
#endif
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              tc::mma_f16_ts(tCagg, tZ + 8 * ks, dWc + dstep + (uint64_t)(hz * (FusedPack6::WC_BYTES / 32) + ks * 64), id32h,
                             hz > 0 || ks > 0);
            if (hz == 0) {
              tc::mma_commit(&ws.bar[3]);
            } else {  // gate GEMM right behind GEMM1b: 0.5 ([h | 1] . [Wr_h | Wz_h ; br | bz] + agg . [Wr_a | Wz_a])
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCzr, tAh + 8 * ks, dBzrh + dstep + (uint64_t)(ks * 128), id64h, ks > 0);
              tc::mma_f16_ts(tCzr, tOnes, dBBzr + dstep, id64h, true);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(tCzr, tCagg + 8 * ks, dBzra + dstep + (uint64_t)(ks * 128), id64t, true);
              tc::mma_commit(&ws.bar[1]);
            }
          }
          __syncwarp();
        }
      }
      F6_PROF(8);
      tc::mbar_wait(&ws.bar[1], ph);
      tc::fence_after_thread_sync();
      F6_PROF(9);
      {  // reset gate -> r * h operand (over the h operand: the gate GEMM has read it)
        float v[32];
        tc::tmem_ld32(tCzr + lane_off, v);
        uint32_t rr[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          rr[i] = tc::pack_f16x2(fz_sigmoid_half<PRECISE>(v[2 * i]) * h[2 * i], fz_sigmoid_half<PRECISE>(v[2 * i + 1]) * h[2 * i + 1]);
        tc::tmem_st16(tAh + lane_off, rr);
      }
      F6_PROF(10);
      operands_ready(2);
      F6_PROF(11);
      // ------------------------------------------------------------ candidate GEMM: [r*h | 1] . [Wh_h ; bh] + agg . Wh_a
      if (mma_warp) {
#if F8_ARRIVE
        tc::mbar_wait(&ws.obar[2], ph);
#endif
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCht, tAh + 8 * ks, dBhh + dstep + (uint64_t)(ks * 64), id32h, ks > 0);
          tc::mma_f16_ts(tCht, tOnes, dBBh + dstep, id32h, true);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(tCht, tCagg + 8 * ks, dBha + dstep + (uint64_t)(ks * 64), id32t, true);
          tc::mma_commit(&ws.bar[2]);
        }
        __syncwarp();
      }
      float z[D];
      {  // update gate, while the candidate GEMM runs (it writes columns [0,32), z's pre-activation is in [32,64))
        float v[32];
        tc::tmem_ld32(tCzr + D + lane_off, v);
#pragma unroll
        for (int c = 0; c < D; ++c) z[c] = fz_sigmoid_half<PRECISE>(v[c]);
      }
      F6_PROF(12);
      tc::mbar_wait(&ws.bar[2], ph);
      tc::fence_after_thread_sync();
      F6_PROF(13);
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
        for (int j = 0; j < D; ++j) h[j] = fmaf(fmaf(gq[j], inv, ninv), gb[j], h[j]) + gb[D + j];
        if (s + 1 < a.steps) {  // packed once: the gathers' shared-memory copy and the next step's GEMM operand
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = tc::pack_f16x2(h[2 * i], h[2 * i + 1]);
#pragma unroll
          for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          tc::tmem_st16(tAh + lane_off, pk);
          tc::tmem_wait_st();
        }
      }
      tc::fence_before_thread_sync();
      F6_PROF(14);
      tc::named_bar_sync(bar_id, F8_THREADS);
      ph ^= 1;
      F6_PROF(15);
    }
    // ---------------------------------------------------------------- GlobalSumPool, 16 columns at a time
    // (the h copy is dead after the last step: its rows take 16 fp32 columns; a half-warp sums one molecule's natural rows)
    {
      const float keep = aid > 0 ? 1.f : 0.f;  // id 0 is not pooled (models/layers.py:163)
      const float* hfp = reinterpret_cast<const float*>(ws.hb);
      const int nm = tp.nm, hl = lane & 15;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          reinterpret_cast<float4*>(hbrow)[c] = make_float4(keep * h[16 * cb + 4 * c], keep * h[16 * cb + 4 * c + 1],
                                                            keep * h[16 * cb + 4 * c + 2], keep * h[16 * cb + 4 * c + 3]);
        tc::named_bar_sync(bar_id, F8_THREADS);
        for (int mi = (t >> 4); mi < nm; mi += 8) {
          const int lo = tp.mol_lo[mi], hi = tp.mol_lo[mi + 1];
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four interleaved partial sums, combined in a fixed order
          int rr = lo;
          for (; rr + 4 <= hi; rr += 4) {
            s0 += hfp[rr * F8_HS + hl], s1 += hfp[(rr + 1) * F8_HS + hl];
            s2 += hfp[(rr + 2) * F8_HS + hl], s3 += hfp[(rr + 3) * F8_HS + hl];
          }
          for (; rr < hi; ++rr) s0 += hfp[rr * F8_HS + hl];
          a.pooled[(size_t)tp.molid[mi] * D + 16 * cb + hl] = (s0 + s1) + (s2 + s3);
        }
        tc::named_bar_sync(bar_id, F8_THREADS);
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  F6_PROF_FLUSH;
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

}  // namespace imp

#ifdef F6_PHASE_PROF
extern "C" void imp_debug_f8_prof(unsigned long long* host_out) {  // reads and clears the phase counters (profiling builds only)
  cudaMemcpyFromSymbol(host_out, imp::f6_prof_total, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {};
  cudaMemcpyToSymbol(imp::f6_prof_total, z, sizeof(z));
}
#endif


using namespace imp;

static int fused8_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

namespace imp {
int launch_fused_h8(const void* d_plan, int32_t n_atoms, int32_t n_cat_atoms, int32_t bond_vocab, const float* d_atom_emb,
                    int32_t atom_vocab, const float* d_bond_emb, int32_t steps, const void* d_packed, float eps, bool precise,
                    float* d_pooled, cudaStream_t st) {
  Fused8Args a;
  a.plan = (const unsigned char*)d_plan, a.atom_emb = d_atom_emb, a.bond_emb = d_bond_emb, a.packed = (const unsigned char*)d_packed;
  a.pooled = d_pooled, a.atom_vocab = atom_vocab, a.bond_vocab = bond_vocab, a.steps = steps, a.eps = eps;
  // one persistent CTA per SM; CTAs are split between the towers in proportion to their atoms
  const int sms = fused8_sm_count();
  int nc = (int)((int64_t)sms * n_cat_atoms / (n_atoms > 0 ? n_atoms : 1));
  nc = nc < 1 ? 1 : (nc > sms - 1 ? sms - 1 : nc);
  int na = sms - nc;
  const int want = (int)ceil_div(ceil_div((int64_t)n_atoms, 100), F8_CTX) + 1;  // never more CTAs than a small batch has tiles for
  if (nc > want) nc = want;
  if (na > want) na = want;
  a.n_cta_cat = nc;
  const size_t smem = (size_t)fused8_smem_bytes(steps, bond_vocab);
  IMP_REQUIRE(smem <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused_planned: needs %zu B of shared memory", smem);
  if (precise) {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h8_kernel<true><<<nc + na, F8_CTX * F8_THREADS, smem, st>>>(a);
  } else {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h8_kernel<false><<<nc + na, F8_CTX * F8_THREADS, smem, st>>>(a);
  }
  IMP_LAUNCH_CHECK();
  return 0;
}
}  // namespace imp
