// Seventh generation of the fused whole-tower forward ("h7"): the step pipeline of generation 6 (fused_fwd6.cu: Z halves ->
// GEMM1 -> gate GEMMs on the in-place tf32 accumulator -> candidate GEMM -> LayerNorm) on EIGHT warps per 128-row tile.
//
// Replaces the same reference code: Embedding -> [BondMatrixMessage o Reduce -> GatedUpdate] x S -> GlobalSumPool
// (train_viscosity.py:163-187, models/layers.py:57-164).
//
// Generation 6 is latency-bound (ncu: 46 % of the issue slots, 0.7 eligible warps per scheduler): one thread owns one atom
// row with all 32 state columns (128 registers), so the register file holds four warps per sub-partition, and inside a tile
// the warp that owns the highest in-degrees finishes its Z rows last while the three others wait at the operand barrier
// (21 % of all warp samples sit there).  Here
//   * a tile (context) is 256 threads: warps q and q + 4 share TMEM quadrant q.  In the gate phases a thread owns one row
//     and 16 of its 32 columns (the LayerNorm statistics cross the warp pair through shared memory and a 64-thread barrier);
//     80 registers per thread, three contexts = 24 warps per SM instead of 16, each with half as long a dependent chain;
//   * the Z rows are built FOUR LANES WIDE: lane 4 g + j accumulates, for the rows in TMEM lanes g and g + 8 of its 16-lane
//     half quadrant, the state columns 8 j .. 8 j + 7 (one 16-byte shared-memory read per entry) against the four bond
//     components of the K half, and the warp writes both rows' fragments with ONE tcgen05.st.16x256b (the mma.sync
//     accumulator layout).  In-degree granularity is 8 rows (an octet) instead of 32, and the plan's in-degree order is
//     dealt out so that every warp gets octets o and 15 - o: all eight warps of a tile carry the same number of entries;
//   * the entries' coefficient vectors mult * c[bond] are formed once per tile, not once per step and K half.
// The K order of a Wc half follows the fragment layout (imp_fused_pack_planned7): K = 16 (m % 8) + 4 (m / 8) + k.
//
// TMEM columns of a context (base = 128 * ctx): as in generation 6
//   [  0, 64)  Z half (A of GEMM1)  ->  r | z pre-activations (D of the gate GEMM)  ->  [0,32) candidate (D of GEMM3)
//   [ 64, 96)  aggregated messages: D of GEMM1 (fp32) = tf32 A operand of the gate and candidate GEMMs
//   [ 96,112)  h operand (16-bit pairs), then r*h operand        [112,120)  the constant (1, 0, ...) bias K-step
#include "fused_common.cuh"
#include "fused_pack6.cuh"
#include "fused_plan.cuh"

namespace imp {

#ifndef F7_CTX
#define F7_CTX 3
#endif
constexpr int F7_THREADS = 256;
constexpr int F7_HS = 16;  // words per row of the shared-memory h copy (64 B: the four lanes of a row read it as 4 x 16 B)

struct alignas(128) FusedWgSmem7 {
  uint32_t hb[FZ_ROWS * F7_HS];  // h as packed halves (natural row order); during the pooling: 16 fp32 columns of h
  // One 32-byte item per entry of the tile, formed once per tile: for each K half the four coefficient halves
  // mult * c[bond][4 hz .. 4 hz + 4) and the shared-memory address of the source row's h copy, so that the Z loop reads ONE
  // 16-byte word per (entry, K half).  item[FP_ECAP] is the all-zero item that pads a row up to its octet's trip count.
  uint4 item[FP_ECAP + 1][2];
  float2 xs[2][FZ_ROWS];         // LayerNorm partial sums (sum, sum of squares) of the two column halves of a row
  FusedTile plan[2];
  uint64_t bar[4];   // 1: gate GEMM done, 2: candidate GEMM done, 3: GEMM1a done
  uint64_t obar[3];  // operands in tensor memory (one arrival per warp): Z half 0, Z half 1, r * h
  uint64_t pbar[2];  // plan buffers
  uint64_t pad[7];
};

__host__ __device__ inline int fused7_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack6::BYTES + ctab + F7_CTX * (int)sizeof(FusedWgSmem7) + (int)sizeof(FusedCtl);
}

struct Fused7Args {
  const unsigned char* plan;
  const float* atom_emb;
  const float* bond_emb;
  const unsigned char* packed;  // [2][steps][FusedPack6::BYTES], Wc in the generation-7 K order
  float* pooled;                // [2P][32]
  int atom_vocab, bond_vocab, steps, n_cta_cat;
  float eps;
};

struct F7True { static constexpr bool value = true; };
struct F7False { static constexpr bool value = false; };

template <bool PRECISE>
__global__ void __launch_bounds__(F7_CTX * F7_THREADS, 1) mpnn_fused_h7_kernel(const Fused7Args a) {
  constexpr int D = FZ_D, DH = FZ_D / 2;
  constexpr int NT = F7_CTX * F7_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 8, t = tid & 255;
  const int q = warp & 3;          // TMEM quadrant of this warp (hardware rule: warp id % 4)
  const int hf = (warp >> 2) & 1;  // gate phases: column half; Z phase: 16-lane half of the quadrant
  const int wbytes = a.steps * FusedPack6::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);
  FusedWgSmem7& ws = reinterpret_cast<FusedWgSmem7*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + F7_CTX * sizeof(FusedWgSmem7));

  const int tower = blockIdx.x >= a.n_cta_cat;
  const FusedPlanHeader* hdr = reinterpret_cast<const FusedPlanHeader*>(a.plan);
  if (__ldg(&hdr->status) == 2) return;  // the plan ran out of tile records: some are unwritten (the host raises, model.check_status)
  const int n_tiles = min(__ldg(&hdr->n_tiles[tower]), __ldg(&hdr->cap[tower]));
  const FusedTile* tiles = reinterpret_cast<const FusedTile*>(a.plan + FP_HEADER_BYTES) + (size_t)(tower ? __ldg(&hdr->cap[0]) : 0);
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const int first = cta_in_tower * F7_CTX + ctx, stride = n_cta_tower * F7_CTX;

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
    for (int i = 0; i < 3; ++i) tc::mbar_init(&ws.obar[i], F7_THREADS / 32);
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
  if (t < 2) ws.item[FP_ECAP][t] = make_uint4(0u, 0u, tc::smem_u32(ws.hb), 0u);
  if (warp == 0) tc::tmem_alloc<512>(&ctl.tmem_base);
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  tc::mbar_wait(&ctl.wbar, 0);

  const uint32_t sw0 = tc::smem_u32(smem);
  const uint32_t tbase = ctl.tmem_base + (uint32_t)(ctx * 128);
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;             // 32x32b accesses: this thread = TMEM lane 32 q + lane
  const uint32_t zlane_off = (uint32_t)(q * 32 + hf * 16) << 16;  // 16x256b store of the Z fragments
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
  const bool mma_warp = (warp & 7) == 0;
  const int bar_id = 1 + ctx;                    // all 256 threads of the context
  const int pair_id = 1 + F7_CTX + ctx * 4 + q;  // the two warps of a quadrant (64 threads)
  static_assert(1 + F7_CTX + 4 * F7_CTX <= 16, "named barriers");

  // Row of this thread in the gate phases: TMEM lane 32 q + lane.  The plan lists the rows by in-degree (ascending); octet o
  // of that order (8 rows) goes to the first 8 lanes of the 16-lane half quadrant zw = o (o < 8), octet 15 - o to its last 8.
  const int zw_of_lane = q + 4 * (lane >> 4);
  const int myslot = 8 * ((lane & 8) ? 15 - zw_of_lane : zw_of_lane) + (lane & 7);
  // Rows of this thread in the Z phase: TMEM lanes 16 hf + g and 16 hf + g + 8 of the quadrant, state columns 8 j .. 8 j + 7
  const int g = lane >> 2, j = lane & 3;
  const int slotA = 8 * (q + 4 * hf) + g, slotB = 8 * (15 - (q + 4 * hf)) + g;
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  const int L = q * 32 + lane;
  const uint32_t hb_s = tc::smem_u32(ws.hb), item_s = tc::smem_u32(ws.item), zero_s = item_s + 32u * FP_ECAP;
  const uint32_t j16 = 16u * (uint32_t)j;
  // "this warp's operand rows are in tensor memory": one arrival per warp; only the MMA-issuing warp waits for all eight, the
  // others go on with their next phase (all warps of a tile carry the same work: what is left is scheduling noise, and a
  // blocking barrier would add every phase's slowest warp to the tile's critical path)
  auto operands_ready = [&](int which) {
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&ws.obar[which]);
  };
  uint32_t ph = 0, pph = 0;  // parities: per-step MMA barriers; plan buffers (bit b = buffer b)
  if (hf == 0) {  // the constant (1, 0, ..., 0) K-step that carries the biases: written once, never overwritten
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
    tc::mbar_wait(&ws.pbar[buf], (pph >> buf) & 1u);
    pph ^= 1u << buf;
    const FusedTile& tp = ws.plan[buf];
    const uint32_t sw = tp.slot[myslot];
    const int r = sw & 127, aid = (int)(sw >> 22);
    const uint32_t swA = tp.slot[slotA], swB = tp.slot[slotB];
    const int degA = (swA >> 7) & 31, degB = (swB >> 7) & 31;
    const int dmaxA = __reduce_max_sync(0xffffffffu, degA), dmaxB = __reduce_max_sync(0xffffffffu, degB);  // octet maxima
    // first item of the two rows (the zero item for a row without entries)
    const uint32_t itA = degA > 0 ? item_s + 32u * ((swA >> 12) & 1023u) : zero_s;
    const uint32_t itB = degB > 0 ? item_s + 32u * ((swB >> 12) & 1023u) : zero_s;
    uint32_t* hbrow = &ws.hb[r * F7_HS + 8 * hf];
    float h[DH];
    {  // Embedding(atom): fp32 state in registers; packed once for the shared-memory copy (gathers) and the GEMM operand
      const float4* er = emb4 + aid * (D / 4) + 4 * hf;
      uint32_t pk[8];
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 x = __ldg(er + c);
        h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
        pk[2 * c] = tc::pack_f16x2(x.x, x.y), pk[2 * c + 1] = tc::pack_f16x2(x.z, x.w);
      }
      reinterpret_cast<uint4*>(hbrow)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      reinterpret_cast<uint4*>(hbrow)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      tc::tmem_st8(tAh + lane_off + (uint32_t)(8 * hf), pk);
    }
    for (int e = t; e < (int)tp.n_ent; e += F7_THREADS) {  // the items, once per tile (generation 6 decodes per step and K half)
      const uint32_t ec = tp.ent[e];
      const uint4 c = s_ctab[(ec >> 8) & 0xff];
      const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
      const __half2 mult = *reinterpret_cast<const __half2*>(&mbits);
      __half2 p0 = __hmul2(*reinterpret_cast<const __half2*>(&c.x), mult), p1 = __hmul2(*reinterpret_cast<const __half2*>(&c.y), mult);
      __half2 p2 = __hmul2(*reinterpret_cast<const __half2*>(&c.z), mult), p3 = __hmul2(*reinterpret_cast<const __half2*>(&c.w), mult);
      const uint32_t haddr = hb_s + (ec & 0x7fu) * (F7_HS * 4);
      ws.item[e][0] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), haddr, 0u);
      ws.item[e][1] = make_uint4(*reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3), haddr, 0u);
    }
    tc::tmem_wait_st();
    tc::fence_before_thread_sync();
    tc::named_bar_sync(bar_id, F7_THREADS);

    for (int s = 0; s < a.steps; ++s) {
      const uint64_t dstep = (uint64_t)(s * (FusedPack6::BYTES / 16));
      const float* gb = reinterpret_cast<const float*>(smem + s * FusedPack6::BYTES + FusedPack6::OFF_BIAS);
      // ------------------------------------------------------------ Z in two K halves -> TMEM -> GEMM1 (-> gate GEMM)
#pragma unroll 1
      for (int hz = 0; hz < 2; ++hz) {
        __half2 accA[16], accB[16];  // [2 c + i]: state column 8 j + c, bond components 4 hz + 2 i, 4 hz + 2 i + 1
        // one item of one row: acc (+)= h[src][8 j .. 8 j + 8) (x) (mult * c[4 hz .. 4 hz + 4))
        auto fma_item = [&](const uint4 it, __half2(&acc)[16], auto first_entry) {
          const __half2 c0 = *reinterpret_cast<const __half2*>(&it.x), c1 = *reinterpret_cast<const __half2*>(&it.y);
          const uint4 hv = tc::lds128(it.z + j16);
          const __half2 hw[4] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y),
                                 *reinterpret_cast<const __half2*>(&hv.z), *reinterpret_cast<const __half2*>(&hv.w)};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half2 lo = __low2half2(hw[i]), hi = __high2half2(hw[i]);
            if constexpr (decltype(first_entry)::value) {
              acc[4 * i] = __hmul2(lo, c0), acc[4 * i + 1] = __hmul2(lo, c1);
              acc[4 * i + 2] = __hmul2(hi, c0), acc[4 * i + 3] = __hmul2(hi, c1);
            } else {
              acc[4 * i] = __hfma2(lo, c0, acc[4 * i]), acc[4 * i + 1] = __hfma2(lo, c1, acc[4 * i + 1]);
              acc[4 * i + 2] = __hfma2(hi, c0, acc[4 * i + 2]), acc[4 * i + 3] = __hfma2(hi, c1, acc[4 * i + 3]);
            }
          }
        };
        // one row: the octet's trip count is warp-uniform; a row with fewer entries runs the rest on the zero item
        auto zrow = [&](uint32_t p, int deg, int dmax, __half2(&acc)[16]) {
          const uint32_t zp = zero_s + 16u * (uint32_t)hz;
          p += 16u * (uint32_t)hz;
          uint4 it = tc::lds128(p);
          {
            p = 1 < deg ? p + 32u : zp;
            const uint4 nx = tc::lds128(p);  // the next item is in flight during this one's FMAs
            fma_item(it, acc, F7True{});
            it = nx;
          }
#pragma unroll 1
          for (int e = 1; e < dmax; ++e) {
            p = e + 1 < deg ? p + 32u : zp;
            const uint4 nx = tc::lds128(p);
            fma_item(it, acc, F7False{});
            it = nx;
          }
        };
        zrow(itA, degA, dmaxA, accA);
        zrow(itB, degB, dmaxB, accB);
        if (hz == 1) {  // GEMM1a must have consumed the first half before its columns are rewritten
          tc::mbar_wait(&ws.bar[3], ph);
          tc::fence_after_thread_sync();
        }
        {
          uint32_t rr[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            rr[4 * c] = *reinterpret_cast<const uint32_t*>(&accA[2 * c]), rr[4 * c + 1] = *reinterpret_cast<const uint32_t*>(&accA[2 * c + 1]);
            rr[4 * c + 2] = *reinterpret_cast<const uint32_t*>(&accB[2 * c]), rr[4 * c + 3] = *reinterpret_cast<const uint32_t*>(&accB[2 * c + 1]);
          }
          tc::tmem_st_16x256b_x8(tZ + zlane_off, rr);
        }
        operands_ready(hz);
        if (mma_warp) {
          tc::mbar_wait(&ws.obar[hz], ph);
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
      tc::mbar_wait(&ws.bar[1], ph);
      tc::fence_after_thread_sync();
      {  // reset gate -> r * h operand (over the h operand: the gate GEMM has read it)
        float v[DH];
        tc::tmem_ld16(tCzr + lane_off + (uint32_t)(DH * hf), v);
        uint32_t rr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rr[i] = tc::pack_f16x2(fz_sigmoid_half<PRECISE>(v[2 * i]) * h[2 * i], fz_sigmoid_half<PRECISE>(v[2 * i + 1]) * h[2 * i + 1]);
        tc::tmem_st8(tAh + lane_off + (uint32_t)(8 * hf), rr);
      }
      operands_ready(2);
      // ------------------------------------------------------------ candidate GEMM: [r*h | 1] . [Wh_h ; bh] + agg . Wh_a
      if (mma_warp) {
        tc::mbar_wait(&ws.obar[2], ph);
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
      float z[DH];
      {  // update gate, while the candidate GEMM runs (it writes columns [0,32), z's pre-activation is in [32,64))
        float v[DH];
        tc::tmem_ld16(tCzr + lane_off + (uint32_t)(D + DH * hf), v);
#pragma unroll
        for (int c = 0; c < DH; ++c) z[c] = fz_sigmoid_half<PRECISE>(v[c]);
      }
      tc::mbar_wait(&ws.bar[2], ph);
      tc::fence_after_thread_sync();
      {  // candidate, blend, LayerNorm (biased variance, eps), residual  (models/layers.py:151-156)
        float gq[DH];
        tc::tmem_ld16(tCht + lane_off + (uint32_t)(DH * hf), gq);
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int c = 0; c < DH; c += 2) {
          const float n0 = fmaf(z[c], fz_tanh<PRECISE>(gq[c]) - h[c], h[c]);
          const float n1 = fmaf(z[c + 1], fz_tanh<PRECISE>(gq[c + 1]) - h[c + 1], h[c + 1]);
          gq[c] = n0, gq[c + 1] = n1;
          s0 += n0, s1 += n1;
          q0 = fmaf(n0, n0, q0), q1 = fmaf(n1, n1, q1);
        }
        ws.xs[hf][L] = make_float2(s0 + s1, q0 + q1);  // the other half of the row's columns lives in the partner warp
        tc::named_bar_sync(pair_id, 64);
        const float2 x0 = ws.xs[0][L], x1 = ws.xs[1][L];
        const float mean = (x0.x + x1.x) * (1.0f / D);
        const float var = fmaxf(fmaf(x0.y + x1.y, 1.0f / D, -mean * mean), 0.f);  // biased variance
        const float inv = PRECISE ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
        const float ninv = -mean * inv;
        const float* gam = gb + DH * hf;
#pragma unroll
        for (int c = 0; c < DH; ++c) h[c] = fmaf(fmaf(gq[c], inv, ninv), gam[c], h[c]) + gam[D + c];
        if (s + 1 < a.steps) {  // packed once: the gathers' shared-memory copy and the next step's GEMM operand
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = tc::pack_f16x2(h[2 * i], h[2 * i + 1]);
          reinterpret_cast<uint4*>(hbrow)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          reinterpret_cast<uint4*>(hbrow)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          tc::tmem_st8(tAh + lane_off + (uint32_t)(8 * hf), pk);
          tc::tmem_wait_st();
        }
      }
      tc::fence_before_thread_sync();
      tc::named_bar_sync(bar_id, F7_THREADS);
      ph ^= 1;
    }
    // ---------------------------------------------------------------- GlobalSumPool, 16 columns at a time
    // (the h copy is dead after the last step: its rows take 16 fp32 columns; a half-warp sums one molecule's natural rows)
    {
      const float keep = aid > 0 ? 1.f : 0.f;  // id 0 is not pooled (models/layers.py:163)
      const float* hfp = reinterpret_cast<const float*>(ws.hb);
      const int nm = tp.nm, hl = lane & 15;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        if (hf == cb) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            reinterpret_cast<float4*>(&ws.hb[r * F7_HS])[c] =
                make_float4(keep * h[4 * c], keep * h[4 * c + 1], keep * h[4 * c + 2], keep * h[4 * c + 3]);
        }
        tc::named_bar_sync(bar_id, F7_THREADS);
        for (int mi = (t >> 4); mi < nm; mi += F7_THREADS / 16) {
          const int lo = tp.mol_lo[mi], hi = tp.mol_lo[mi + 1];
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four interleaved partial sums, combined in a fixed order
          int rr = lo;
          for (; rr + 4 <= hi; rr += 4) {
            s0 += hfp[rr * F7_HS + hl], s1 += hfp[(rr + 1) * F7_HS + hl];
            s2 += hfp[(rr + 2) * F7_HS + hl], s3 += hfp[(rr + 3) * F7_HS + hl];
          }
          for (; rr < hi; ++rr) s0 += hfp[rr * F7_HS + hl];
          a.pooled[(size_t)tp.molid[mi] * D + 16 * cb + hl] = (s0 + s1) + (s2 + s3);
        }
        tc::named_bar_sync(bar_id, F7_THREADS);
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

}  // namespace imp

using namespace imp;

static int fused7_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

namespace imp {
int launch_fused_h7(const void* d_plan, int32_t n_atoms, int32_t n_cat_atoms, int32_t bond_vocab, const float* d_atom_emb,
                    int32_t atom_vocab, const float* d_bond_emb, int32_t steps, const void* d_packed, float eps, bool precise,
                    float* d_pooled, cudaStream_t st) {
  Fused7Args a;
  a.plan = (const unsigned char*)d_plan, a.atom_emb = d_atom_emb, a.bond_emb = d_bond_emb, a.packed = (const unsigned char*)d_packed;
  a.pooled = d_pooled, a.atom_vocab = atom_vocab, a.bond_vocab = bond_vocab, a.steps = steps, a.eps = eps;
  // one persistent CTA per SM; CTAs are split between the towers in proportion to their atoms
  const int sms = fused7_sm_count();
  int nc = (int)((int64_t)sms * n_cat_atoms / (n_atoms > 0 ? n_atoms : 1));
  nc = nc < 1 ? 1 : (nc > sms - 1 ? sms - 1 : nc);
  int na = sms - nc;
  const int want = (int)ceil_div(ceil_div((int64_t)n_atoms, 100), F7_CTX) + 1;  // never more CTAs than a small batch has tiles for
  if (nc > want) nc = want;
  if (na > want) na = want;
  a.n_cta_cat = nc;
  const size_t smem = (size_t)fused7_smem_bytes(steps, bond_vocab);
  IMP_REQUIRE(smem <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused_planned: needs %zu B of shared memory", smem);
  if (precise) {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h7_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h7_kernel<true><<<nc + na, F7_CTX * F7_THREADS, smem, st>>>(a);
  } else {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h7_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h7_kernel<false><<<nc + na, F7_CTX * F7_THREADS, smem, st>>>(a);
  }
  IMP_LAUNCH_CHECK();
  return 0;
}
}  // namespace imp
