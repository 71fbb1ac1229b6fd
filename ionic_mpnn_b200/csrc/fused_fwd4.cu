// Fourth generation of the fused whole-tower forward ("h4", the default for IEEE-half operands).
//
// Same mathematics, row ownership and packed-HFMA2 Z build as the third generation (fused_fwd.cu: Embedding ->
// [BondMatrixMessage o Reduce -> GatedUpdate] x S -> GlobalSumPool, train_viscosity.py:163-187, models/layers.py:57-164;
// four 128-thread contexts per CTA, one thread per atom row, 128 TMEM columns per context).  What changed, and why
// (profiles/README.md, round 2: issue slots 50 % busy, 25 % of the stall samples at the context barriers, 10 k warp
// instructions per pair of which the fp32 gate epilogue is the largest part):
//
//   * no blocking barrier inside a step.  After writing its operand rows to tensor memory a warp ARRIVES on a shared
//     counter and moves on; the warp that arrives last issues the MMAs.  A light warp (rows are sorted by in-degree) flows
//     from the first Z half into the second instead of parking, and every warp only ever waits for data it needs (the
//     mbarriers of the MMA groups, and one mbarrier for "all rows of h are rewritten" that is armed at the end of a step
//     and awaited just before the next step's first gather);
//   * TMEM columns are laid out so that nothing but the aggregated messages has to be converted between GEMM1 and
//     GEMM2: [0,64) Z half -> z|r accumulator -> candidate accumulator; [64,96) agg accumulator -> agg and tanh(r')*h
//     operands; [96,112) h operand, written once per step at the end of the previous step (the same packed words go to the
//     shared-memory copy the gathers read); [112,120) the constant (1, 0, ...) bias K-step, written once per kernel;
//   * r * h = 0.5 tanh(y_r) h + 0.5 h: the second term is folded into GEMM3 (h is already a resident operand and the
//     top half of Wh is packed pre-multiplied by 0.5; the same shared-memory block serves both K-steps), so the gate
//     epilogue is one tanh and one multiply per element;
//   * Z accumulators start from the first entry's products (no zero fill + FMA); LayerNorm is 3 instructions per element
//     (t = n * inv - mean * inv; h = t * gamma + h + beta) with four interleaved partial sums for the statistics.
#include "fused_common.cuh"

namespace imp {

constexpr int F4_CTX = 4;
constexpr int F4_THREADS = 128;
constexpr int F4_ECAP = 672;

struct alignas(16) FusedWgSmem4 {
  uint32_t hb[FZ_ROWS * FZ_HS];  // words 0..15 of a row: h as packed halves; after the last step: fp32 h rows for the pooling
  int molp[FZ_GROUP + 4];
  int se0[FZ_ROWS], se1[FZ_ROWS], said[FZ_ROWS];
  int cnt[4][8];
  int mole[FZ_GROUP + 4];
  int wsum[4];
  uint32_t ent[F4_ECAP];
  unsigned char rowof[FZ_ROWS];
  unsigned char amask[FZ_ROWS];
  uint64_t bar[6];        // 0: GEMM1 done, 1: GEMM2 done, 2: GEMM3 done, 3: GEMM1a done, 4: all rows of h rewritten
  uint32_t arrivals;      // operand-ready arrivals of the four warps (monotonic; the warp that makes it 4k issues the MMAs)
  uint32_t pad[3];
};

__host__ __device__ inline int fused4_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack::BYTES + ctab + F4_CTX * (int)sizeof(FusedWgSmem4) + (int)sizeof(FusedCtl);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// One warp reports "my rows of the operand are in tensor memory" (tcgen05.wait::st + fence already executed).  Returns true
// in the warp that arrived last: all four warps' rows are then visible to the MMAs it issues.
__device__ __forceinline__ bool f4_arrive_is_last(uint32_t* counter, int lane) {
  uint32_t old = 0;
  __syncwarp();
  if (lane == 0) asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(tc::smem_u32(counter)) : "memory");
  return __any_sync(0xffffffffu, lane == 0 && (old & 3u) == 3u);
}

// Experiment switches (tools/fused_variants.py builds A/B copies of the library with -DF4_SYNC=1 / -DF4_PEEL=0).
#ifndef F4_SYNC
#define F4_SYNC 0  // 0: arrive-and-continue, the warp that arrives last issues the MMAs; 1: blocking context barrier, warp 0 issues
#endif
#ifndef F4_PEEL
#define F4_PEEL 1  // 1: Z accumulators start from the first entry's products; 0: zero fill + FMA
#endif

struct F4True { static constexpr bool value = true; };
struct F4False { static constexpr bool value = false; };

template <bool PRECISE, bool COMPACT>
__global__ void __launch_bounds__(F4_CTX * F4_THREADS, 1) mpnn_fused_h4_kernel(const FusedArgs a) {
  constexpr int D = FZ_D;
  constexpr int FMT = tc::FMT_F16;
  constexpr int NT = F4_CTX * F4_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 7, t = tid & 127, wq = warp & 3;
  const int wbytes = a.steps * FusedPack::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);
  FusedWgSmem4& ws = reinterpret_cast<FusedWgSmem4*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + F4_CTX * sizeof(FusedWgSmem4));

  const int tower = blockIdx.x >= a.n_cta_cat;
  if (tid == 0) {  // resident weights of this tower (all steps): one TMA bulk copy
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
    tc::mbar_init(&ws.bar[4], F4_THREADS);
    ws.arrivals = 0;
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  tc::mbar_wait(&ctl.wbar, 0);

  const uint32_t sw0 = tc::smem_u32(smem);
  const uint32_t tbase = ctl.tmem_base + (uint32_t)(ctx * 128);
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t tZ = tbase, tCzr = tbase, tCht = tbase;
  const uint32_t tCagg = tbase + 64, tAagg = tbase + 64, tArh = tbase + 80, tAh = tbase + 96, tOnes = tbase + 112;
  const uint32_t idesc32 = tc::make_idesc(FMT, FZ_ROWS, D), idesc64 = tc::make_idesc(FMT, FZ_ROWS, 2 * D);
  const uint64_t dWc = tc::make_smem_desc(sw0, D * 16, 128);
  const uint64_t dBzr = tc::make_smem_desc(sw0 + FusedPack::OFF_BZR, 2 * D * 16, 128);
  const uint64_t dBh = tc::make_smem_desc(sw0 + FusedPack::OFF_BH, D * 16, 128);
  const uint64_t dBBzr = tc::make_smem_desc(sw0 + FusedPack::OFF_BBZR, 2 * D * 16, 128);
  const uint64_t dBBh = tc::make_smem_desc(sw0 + FusedPack::OFF_BBH, D * 16, 128);
  const int bar_id = 1 + ctx;
  const bool descending = ctx & 1;
#if F4_SYNC == 0
#define F4_READY() f4_arrive_is_last(&ws.arrivals, lane)
#else
  auto f4_blocking_ready = [&]() {
    tc::named_bar_sync(bar_id, F4_THREADS);
    return wq == 0;
  };
#define F4_READY() f4_blocking_ready()
#endif
  {  // the constant (1, 0, ..., 0) K-step that carries the biases: written once, never overwritten
    const uint32_t ones[8] = {0x00003c00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tc::tmem_st8(tOnes + lane_off, ones);
    tc::tmem_wait_st();
  }

  const int P = a.n_pairs;
  const int n_groups = (P + FZ_GROUP - 1) / FZ_GROUP;
  const int n_cta_tower = tower ? (int)gridDim.x - a.n_cta_cat : a.n_cta_cat;
  const int cta_in_tower = tower ? (int)blockIdx.x - a.n_cta_cat : (int)blockIdx.x;
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  uint32_t ph = 0, hph = 0;  // parities of the per-step MMA barriers and of the "h rewritten" barrier

  for (int g = cta_in_tower * F4_CTX + ctx; g < n_groups; g += n_cta_tower * F4_CTX) {
    const int m0 = g * FZ_GROUP, nm = min(FZ_GROUP, P - m0);
    const int base_mol = tower * P + m0;
    tc::named_bar_sync(bar_id, F4_THREADS);
    if (t <= nm) {
      ws.molp[t] = __ldg(a.mol_ptr + base_mol + t);
      if (COMPACT) ws.mole[t] = __ldg(a.mol_eptr + base_mol + t);
    }
    tc::named_bar_sync(bar_id, F4_THREADS);
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
          tc::named_bar_sync(bar_id, F4_THREADS);
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
      tc::named_bar_sync(bar_id, F4_THREADS);
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
      tc::named_bar_sync(bar_id, F4_THREADS);
      // ---------------------------------------------------------------- decoded entries of the tile -> shared memory
      const int E0 = ws.se0[0], n_ent = ws.se1[max(rows, 1) - 1] - E0;
      const bool staged = n_ent <= F4_ECAP && a.bond_vocab <= 256;
      if (staged)
        for (int i = t; i < n_ent; i += F4_THREADS) {
          if (COMPACT) {
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
      // ---------------------------------------------------------------- thread t owns row r
      const int r = ws.rowof[t];
      const int e0 = ws.se0[r], e1 = ws.se1[r];
      uint32_t* hbrow = &ws.hb[r * FZ_HS];
      float h[D];
      {  // Embedding(atom): fp32 state in registers, packed copy to shared memory (gathers) and tensor memory (GEMM operand)
        const bool valid = r < rows;
        const int id = min(max(ws.said[r], 0), a.atom_vocab - 1);
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < D / 4; ++c) {
          const float4 x = valid ? __ldg(emb4 + id * (D / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
          pk[2 * c] = tc::pack_f16x2(x.x, x.y), pk[2 * c + 1] = tc::pack_f16x2(x.z, x.w);
        }
#pragma unroll
        for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        tc::tmem_st16(tAh + lane_off, pk);
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
      tc::tmem_wait_st();
      tc::fence_before_thread_sync();
      mbar_arrive(&ws.bar[4]);  // h_0 of this row is in place (shared-memory copy and operand)

      for (int s = 0; s < a.steps; ++s) {
        const uint64_t dstep = (uint64_t)(s * (FusedPack::BYTES / 16));
        const float* bias = reinterpret_cast<const float*>(smem + s * FusedPack::BYTES + FusedPack::OFF_BIAS);
        // every row of h (and, on the first step, the staged entries) must be in place before the first gather
        tc::mbar_wait(&ws.bar[4], hph);
        hph ^= 1;
        // ------------------------------------------------------------ Z in two K halves -> TMEM -> GEMM1
#pragma unroll 1
        for (int hz = 0; hz < 2; ++hz) {
          __half2 acc[D * 2];
          // one entry: acc (+)= h[src] (x) (mult * c[4 hz .. 4 hz + 4))
          auto entry = [&](uint32_t src, uint32_t bond, __half2 mult, auto first) {
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
                if constexpr (decltype(first)::value) {
                  acc[m * 2] = __hmul2(lo, c0), acc[m * 2 + 1] = __hmul2(lo, c1);
                  acc[m * 2 + 2] = __hmul2(hi, c0), acc[m * 2 + 3] = __hmul2(hi, c1);
                } else {
                  acc[m * 2] = __hfma2(lo, c0, acc[m * 2]), acc[m * 2 + 1] = __hfma2(lo, c1, acc[m * 2 + 1]);
                  acc[m * 2 + 2] = __hfma2(hi, c0, acc[m * 2 + 2]), acc[m * 2 + 3] = __hfma2(hi, c1, acc[m * 2 + 3]);
                }
              }
            }
          };
#if !F4_PEEL
          if (staged) {
#pragma unroll
            for (int i = 0; i < D * 2; ++i) acc[i] = __half2(__ushort_as_half(0), __ushort_as_half(0));
            uint32_t ec = 0u, en = e0 < e1 ? ws.ent[e0 - E0] : 0u;
#pragma unroll 1
            for (int e = e0; e < e1; ++e) {
              ec = en;
              if (e + 1 < e1) en = ws.ent[e + 1 - E0];
              const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
              entry(ec & 0xff, (ec >> 8) & 0xff, *reinterpret_cast<const __half2*>(&mbits), F4False{});
            }
          } else
#endif
          if (e0 >= e1) {
#pragma unroll
            for (int i = 0; i < D * 2; ++i) acc[i] = __half2(__ushort_as_half(0), __ushort_as_half(0));
          } else if (staged) {
            uint32_t ec = ws.ent[e0 - E0];
            uint32_t en = e0 + 1 < e1 ? ws.ent[e0 + 1 - E0] : 0u;
            {
              const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
              entry(ec & 0xff, (ec >> 8) & 0xff, *reinterpret_cast<const __half2*>(&mbits), F4True{});
            }
#pragma unroll 1
            for (int e = e0 + 1; e < e1; ++e) {
              ec = en;
              if (e + 1 < e1) en = ws.ent[e + 1 - E0];  // next entry's descriptor is in flight during this one's FMAs
              const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
              entry(ec & 0xff, (ec >> 8) & 0xff, *reinterpret_cast<const __half2*>(&mbits), F4False{});
            }
          } else {
            int mbase = 0;  // COMPACT: first row of the molecule that owns row r (rare path)
            if (COMPACT)
              for (int mi = ms; mi < me; ++mi)
                if (ws.molp[mi] - a0 <= r) mbase = ws.molp[mi] - a0;
            auto decode = [&](int e, uint32_t& src, uint32_t& bond, __half2& mult) {
              int bm, sr;
              if (!COMPACT) {
                bm = __ldg(a.edge_bm + e);
                sr = __ldg(a.col_src + e) - a0;
              } else {
                const uint32_t w = __ldg(a.edge_w + e);
                bm = (int)(((w >> 8) & 0xffu) | ((w >> 16) & 0xffu) << 16);
                sr = (int)(w & 0xffu) + mbase;
              }
              src = (uint32_t)min(max(sr, 0), FZ_ROWS - 1);
              mult = __float2half2_rn((float)(bm >> 16));
              bond = (uint32_t)min(bm & 0xffff, a.bond_vocab - 1);
            };
            uint32_t src, bond;
            __half2 mult;
            decode(e0, src, bond, mult);
            entry(src, bond, mult, F4True{});
#pragma unroll 1
            for (int e = e0 + 1; e < e1; ++e) {
              decode(e, src, bond, mult);
              entry(src, bond, mult, F4False{});
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
          if (F4_READY()) {
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
        {  // aggregated messages as a 16-bit A operand (in place over the first half of their accumulator)
          float v[32];
          tc::tmem_ld32(tCagg + lane_off, v);
          uint32_t rr[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) rr[i] = tc::pack_f16x2(v[2 * i], v[2 * i + 1]);
          tc::tmem_st16(tAagg + lane_off, rr);
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        // ------------------------------------------------------------ GEMM2: 0.5 ([h | agg | 1] . [Wz | Wr ; bz | br])
        if (F4_READY()) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCzr, tAh + 8 * ks, dBzr + dstep + (uint64_t)(ks * 128), idesc64, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCzr, tAagg + 8 * ks, dBzr + dstep + (uint64_t)((ks + 2) * 128), idesc64, true);
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
          for (int i = 0; i < 16; ++i)  // tanh(y_r) * h: the 0.5 and the "+ 0.5 h" of r * h ride in GEMM3
            rr[i] = tc::pack_f16x2(fz_tanh<PRECISE>(v[2 * i]) * h[2 * i], fz_tanh<PRECISE>(v[2 * i + 1]) * h[2 * i + 1]);
          tc::tmem_st16(tArh + lane_off, rr);
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        // ------------------------------------------------------------ GEMM3: [agg | tanh(y_r) h | h | 1] . [Wh[d:2d] ; Wh[0:d]/2 ; Wh[0:d]/2 ; bh]
        if (F4_READY()) {
          tc::fence_after_thread_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCht, tAagg + 8 * ks, dBh + dstep + (uint64_t)((ks + 2) * 64), idesc32, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCht, tArh + 8 * ks, dBh + dstep + (uint64_t)(ks * 64), idesc32, true);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) tc::mma_f16_ts(tCht, tAh + 8 * ks, dBh + dstep + (uint64_t)(ks * 64), idesc32, true);
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
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
          for (int j = 0; j < D; j += 4) {
            const float n0 = fmaf(z[j], fz_tanh<PRECISE>(gq[j]) - h[j], h[j]);
            const float n1 = fmaf(z[j + 1], fz_tanh<PRECISE>(gq[j + 1]) - h[j + 1], h[j + 1]);
            const float n2 = fmaf(z[j + 2], fz_tanh<PRECISE>(gq[j + 2]) - h[j + 2], h[j + 2]);
            const float n3 = fmaf(z[j + 3], fz_tanh<PRECISE>(gq[j + 3]) - h[j + 3], h[j + 3]);
            gq[j] = n0, gq[j + 1] = n1, gq[j + 2] = n2, gq[j + 3] = n3;
            s0 += n0, s1 += n1, s2 += n2, s3 += n3;
            q0 = fmaf(n0, n0, q0), q1 = fmaf(n1, n1, q1), q2 = fmaf(n2, n2, q2), q3 = fmaf(n3, n3, q3);
          }
          const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / D);
          const float var = fmaxf(fmaf((q0 + q1) + (q2 + q3), 1.0f / D, -mean * mean), 0.f);  // biased variance
          const float inv = PRECISE ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
          const float ninv = -mean * inv;
#pragma unroll
          for (int j = 0; j < D; ++j) h[j] = fmaf(fmaf(gq[j], inv, ninv), bias[3 * D + j], h[j]) + bias[4 * D + j];
          if (s + 1 < a.steps) {  // packed once: the gathers' shared-memory copy and the next step's GEMM operand
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = tc::pack_f16x2(h[2 * i], h[2 * i + 1]);
#pragma unroll
            for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            tc::tmem_st16(tAh + lane_off, pk);
            tc::tmem_wait_st();
            tc::fence_before_thread_sync();
            mbar_arrive(&ws.bar[4]);
          } else {  // last step: fp32 rows for the pooling (every gather of this tile is done)
#pragma unroll
            for (int c = 0; c < D / 4; ++c)
              reinterpret_cast<float4*>(hbrow)[c] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
          }
        }
        ph ^= 1;
      }
      tc::fence_before_thread_sync();
      tc::named_bar_sync(bar_id, F4_THREADS);
      // ---------------------------------------------------------------- GlobalSumPool
      {
        const float* hfp = reinterpret_cast<const float*>(ws.hb);
        for (int mi = ms + (t >> 5); mi < me; mi += 4) {
          const int lo = ws.molp[mi] - a0, hi = min(ws.molp[mi + 1] - a0, FZ_ROWS);
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
      tc::named_bar_sync(bar_id, F4_THREADS);
      ms = me;
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(ctl.tmem_base);
}

// launch (called from fused_forward_impl in fused_fwd.cu)
int launch_fused_h4(const FusedArgs& a, int grid, bool precise, bool compact, cudaStream_t st) {
  const size_t smem = (size_t)fused4_smem_bytes(a.steps, a.bond_vocab);
  IMP_REQUIRE(smem <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused: needs %zu B of shared memory", smem);
#define IMP_LAUNCH_H4(PREC, CMP)                                                                                              \
  do {                                                                                                                        \
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h4_kernel<PREC, CMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mpnn_fused_h4_kernel<PREC, CMP><<<grid, F4_CTX * F4_THREADS, smem, st>>>(a);                                             \
  } while (0)
  if (precise) {
    if (compact) IMP_LAUNCH_H4(true, true); else IMP_LAUNCH_H4(true, false);
  } else {
    if (compact) IMP_LAUNCH_H4(false, true); else IMP_LAUNCH_H4(false, false);
  }
#undef IMP_LAUNCH_H4
#undef F4_READY
  IMP_LAUNCH_CHECK();
  return 0;
}

}  // namespace imp
