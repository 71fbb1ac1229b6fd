// Sixth generation of the fused whole-tower forward ("h6", the default of the planned forward): generation 5 (tile plan,
// one TMA bulk copy per tile; fused_fwd5.cu) with ONE tcgen05 round trip less per step.
//
// Replaces the same reference code: Embedding -> [BondMatrixMessage o Reduce -> GatedUpdate] x S -> GlobalSumPool
// (train_viscosity.py:163-187, models/layers.py:57-164).
//
// In generations 1-5 the aggregated messages leave GEMM1 as an fp32 accumulator in tensor memory, are read back by the
// row's thread, rounded to 16 bits and written to tensor memory again as the A operand of the gate GEMMs -- a
// tcgen05.commit / mbarrier wait, 32 + 16 registers of tcgen05.ld / st traffic, a context barrier and a second MMA issue
// per step.  Here the gate GEMMs read the accumulator IN PLACE: a kind::tf32 tcgen05.mma takes its A operand from tensor
// memory as one 32-bit element per column, which is exactly the accumulator layout, so
//     [z | r]  = 0.5 ([h | 1] . [Wz_h | Wr_h ; bz | br])  (kind::f16)  +  0.5 agg . [Wz_a | Wr_a]   (kind::tf32)
//     cand     =      [r*h | 1] . [Wh_h ; bh]              (kind::f16)  +      agg . Wh_a            (kind::tf32)
// accumulate into the same fp32 columns, and GEMM1b and the gate GEMM are issued back to back by the same thread
// (tcgen05.mma of one thread execute in issue order).  tf32 keeps 10 explicit significand bits, as IEEE half does, with
// fp32 range.  The update gate z is evaluated while the candidate GEMM runs (its pre-activation columns are not the ones
// that GEMM overwrites), and h is packed to 16 bits once per step for both its shared-memory copy and its operand.
//
// TMEM columns of a context (base = 128 * ctx):
//   [  0, 64)  Z half (A of GEMM1)  ->  r | z pre-activations (D of the gate GEMM)  ->  [0,32) candidate (D of GEMM3)
//   [ 64, 96)  aggregated messages: D of GEMM1 (fp32) = tf32 A operand of the gate and candidate GEMMs
//   [ 96,112)  h operand (16-bit pairs), then r*h operand        [112,120)  the constant (1, 0, ...) bias K-step
#include "fused_common.cuh"
#include "fused_pack6.cuh"
#include "fused_plan.cuh"
#include "fused_prof.cuh"

namespace imp {

constexpr int F6_CTX = 4;
constexpr int F6_THREADS = 128;
// Experiment switch (tools/fused_variants.py): evaluate the r and z gates with tanh.approx.f16x2 (one MUFU per two elements).
// Measured on B200: 2 % SLOWER (8.30 vs 8.13 ms per 524 288 pairs) and less accurate -- the conversions and half-rate HFMA2
// cost more issue slots than the saved MUFU time is worth -- so it stays off.
#ifndef F6_PACKED_GATES
#define F6_PACKED_GATES 0
#endif
// Experiment switch: 1 = the warps that do not issue MMAs only ARRIVE at the operand-ready barriers (bar.arrive) and move on;
// the issuing warp -- the one that owns the highest in-degrees and therefore finishes its Z rows last -- waits (bar.sync).
#ifndef F6_ARRIVE
#define F6_ARRIVE 0  // measured on B200: 1 is 1.5 % slower (9.54 vs 9.40 ms per 524 288 pairs)
#endif
constexpr int F6_HS = 20;  // words per row of the shared-memory h copy (80 B: 16 halves pairs + pad, conflict-free 16-byte reads)

// Gate order inside the 64-wide block: n < 32 is the reset gate r, n >= 32 the update gate z.
// zperm7: K order of a Wc half for the seventh generation (fused_fwd7.cu), whose Z rows are built four lanes wide and stored
// with tcgen05.st.16x256b.
__global__ void fused_pack6_kernel(const float* __restrict__ W /* [K, d, d] */, imp_gru_weights_t w, unsigned char* __restrict__ out,
                                   int zperm7) {
  constexpr int D = FZ_D, KK = FZ_D * FZ_K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D * KK) {  // Wc_hz[l][K(m, k)] = W[k][l][m]  (models/layers.py:108 re-associated, see fused_fwd.cu)
    const int l = i / KK, kk = i % KK, m = kk / FZ_K, k = kk % FZ_K;
    // K halves split the state columns m (16 each); inside a half K = 8 (m % 16) + k
    // (seventh generation: state column m = 16 hz + 4 j + c  ->  K = 16 (2 c + k / 4) + 4 j + k % 4, the 16x256b fragment order)
    const int hz = m / 16, ml = m % 16;
    const int kq = zperm7 ? 16 * (2 * (ml % 4) + k / 4) + 4 * (ml / 4) + (k % 4) : ml * 8 + k;
    *reinterpret_cast<uint16_t*>(out + hz * (FusedPack6::WC_BYTES / 2) + tc::chunk_off(l, kq / 8, D) + (kq % 8) * 2) =
        tc::cvt16<tc::FMT_F16>(W[(k * D + l) * D + m]);
  }
  if (i < 2 * D * D) {  // rows k < d of Wr / Wz multiply h (models/layers.py:146-148: concat([h, agg]))
    const int n = i / D, k = i % D;
    const float vh = 0.5f * (n < D ? w.Wr[k * D + n] : w.Wz[k * D + (n - D)]);
    const float va = 0.5f * (n < D ? w.Wr[(D + k) * D + n] : w.Wz[(D + k) * D + (n - D)]);
    *reinterpret_cast<uint16_t*>(out + FusedPack6::OFF_BZRH + tc::chunk_off(n, k / 8, 2 * D) + (k % 8) * 2) = tc::cvt16<tc::FMT_F16>(vh);
    *reinterpret_cast<uint32_t*>(out + FusedPack6::OFF_BZRA + tc::chunk_off(n, k / 4, 2 * D) + (k % 4) * 4) = f32_to_tf32(va);
  }
  if (i < D * D) {  // Wh: rows k < d multiply r * h, rows d + k multiply agg (models/layers.py:150-151)
    const int n = i / D, k = i % D;
    *reinterpret_cast<uint16_t*>(out + FusedPack6::OFF_BHH + tc::chunk_off(n, k / 8, D) + (k % 8) * 2) = tc::cvt16<tc::FMT_F16>(w.Wh[k * D + n]);
    *reinterpret_cast<uint32_t*>(out + FusedPack6::OFF_BHA + tc::chunk_off(n, k / 4, D) + (k % 4) * 4) = f32_to_tf32(w.Wh[(D + k) * D + n]);
  }
  if (i < 2 * D * 16) {  // bias block of the gates: element (n, k) of a [64 x 16] K-major tile, only k = 0 is non-zero
    const int n = i / 16, k = i % 16;
    const float v = k == 0 ? 0.5f * (n < D ? w.br[n] : w.bz[n - D]) : 0.f;
    *reinterpret_cast<uint16_t*>(out + FusedPack6::OFF_BBZR + tc::chunk_off(n, k / 8, 2 * D) + (k % 8) * 2) = tc::cvt16<tc::FMT_F16>(v);
  }
  if (i < D * 16) {  // bias block of the candidate
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<uint16_t*>(out + FusedPack6::OFF_BBH + tc::chunk_off(n, k / 8, D) + (k % 8) * 2) =
        tc::cvt16<tc::FMT_F16>(k == 0 ? w.bh[n] : 0.f);
  }
  if (i < D) {
    float* b = reinterpret_cast<float*>(out + FusedPack6::OFF_BIAS);
    b[i] = w.gamma[i], b[D + i] = w.beta[i];
  }
}

struct alignas(128) FusedWgSmem6 {
  uint32_t hb[FZ_ROWS * F6_HS];  // words 0..15 of a row: h as packed halves; during the pooling: 16 fp32 columns of h
  FusedTile plan[2];
  uint64_t bar[4];   // 1: gate GEMM done, 2: candidate GEMM done, 3: GEMM1a done
  uint64_t pbar[2];  // plan buffers
  int next_tile[2];  // tile index fetched for the next iteration (double-buffered: written by thread 0, read by all)
  uint64_t pad[9];
};

__host__ __device__ inline int fused6_smem_bytes(int steps, int bond_vocab) {
  const int ctab = (bond_vocab * 16 + 127) / 128 * 128;
  return steps * FusedPack6::BYTES + ctab + F6_CTX * (int)sizeof(FusedWgSmem6) + (int)sizeof(FusedCtl);
}

struct Fused6Args {
  const unsigned char* plan;
  const float* atom_emb;
  const float* bond_emb;
  const unsigned char* packed;  // [2][steps][FusedPack6::BYTES]
  float* pooled;                // [2P][32]
  int atom_vocab, bond_vocab, steps, n_cta_cat;
  float eps;
  int* tickets;  // [2] per-tower tile counters (words 5, 6 of the plan header, zeroed by the launcher): dynamic tile scheduling
  int emb_smem;  // 1: the atom-embedding table is staged in shared memory behind the per-context buffers (atom_vocab * 128 bytes)
};

__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

struct F6True { static constexpr bool value = true; };
struct F6False { static constexpr bool value = false; };

template <bool PRECISE>
__global__ void __launch_bounds__(F6_CTX * F6_THREADS, 1) mpnn_fused_h6_kernel(const Fused6Args a) {
  constexpr int D = FZ_D;
  constexpr int NT = F6_CTX * F6_THREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ctx = tid >> 7, t = tid & 127, wq = warp & 3;
  F6_PROF_DECL;
  const int wbytes = a.steps * FusedPack6::BYTES;
  const int ctab_bytes = (a.bond_vocab * 16 + 127) / 128 * 128;
  uint4* s_ctab = reinterpret_cast<uint4*>(smem + wbytes);
  FusedWgSmem6& ws = reinterpret_cast<FusedWgSmem6*>(smem + wbytes + ctab_bytes)[ctx];
  FusedCtl& ctl = *reinterpret_cast<FusedCtl*>(smem + wbytes + ctab_bytes + F6_CTX * sizeof(FusedWgSmem6));
  // Embedding(atom) table, when it fits: 16-byte chunk c of row r at chunk c ^ (r & 7) -- the 32 lanes of a warp read chunk c
  // of 32 different rows at once, which would all fall into one bank group without the swizzle.  (From global memory the same
  // read is 32 L1 lines per instruction, 1 k cycles of L1 tag throughput per tile: 8.18 -> 7.85 ms per 524 288 pairs.)
  float4* s_emb = reinterpret_cast<float4*>(smem + wbytes + ctab_bytes + F6_CTX * sizeof(FusedWgSmem6) + 128);

  const int tower = blockIdx.x >= a.n_cta_cat;
  const FusedPlanHeader* hdr = reinterpret_cast<const FusedPlanHeader*>(a.plan);
  if (__ldg(&hdr->status) == 2) return;  // the plan ran out of tile records: some are unwritten (the host raises, model.check_status)
  const int n_tiles = min(__ldg(&hdr->n_tiles[tower]), __ldg(&hdr->cap[tower]));
  const FusedTile* tiles = reinterpret_cast<const FusedTile*>(a.plan + FP_HEADER_BYTES) + (size_t)(tower ? __ldg(&hdr->cap[0]) : 0);

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
    // Tiles are handed out by a per-tower ticket counter, not by a fixed stride: contexts whose tiles happen to carry more
    // entries no longer finish last (2.6 % of the warp samples of the strided form sat in EXIT)
    const int c0 = atomicAdd(a.tickets + tower, 1);
    ws.next_tile[0] = c0;
    if (c0 < n_tiles) {  // first tile record of this context
      tc::mbar_arrive_expect_tx(&ws.pbar[0], (uint32_t)sizeof(FusedTile));
      tc::bulk_copy_g2s(&ws.plan[0], tiles + c0, (uint32_t)sizeof(FusedTile), &ws.pbar[0]);
    }
  }
  for (int i = tid; i < a.bond_vocab; i += NT) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(a.bond_emb) + 2 * i + 1);
    s_ctab[i] = make_uint4(tc::pack_f16x2(c0.x, c0.y), tc::pack_f16x2(c0.z, c0.w), tc::pack_f16x2(c1.x, c1.y),
                           tc::pack_f16x2(c1.z, c1.w));
  }
  if (a.emb_smem) {
    for (int i = tid; i < a.atom_vocab * 8; i += NT) {
      const int r = i >> 3, c = i & 7;
      s_emb[r * 8 + (c ^ (r & 7))] = __ldg(reinterpret_cast<const float4*>(a.atom_emb) + i);
    }
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
#if F6_ARRIVE
  const bool mma_warp = wq == ((ctx & 1) ? 0 : 3);  // the warp with the highest in-degrees (see myslot)
#else
  const bool mma_warp = wq == 0;
#endif
  const int bar_id = 1 + ctx, opbar_id = 5 + ctx;
  // "my operand rows are in tensor memory": the issuing warp waits for all four warps, the others do not wait
  auto operands_ready = [&]() {
#if F6_ARRIVE
    if (mma_warp) tc::named_bar_sync(opbar_id, F6_THREADS);
    else named_bar_arrive(opbar_id, F6_THREADS);
#else
    tc::named_bar_sync(opbar_id, F6_THREADS);
#endif
  };
  const int myslot = (ctx & 1) ? FZ_ROWS - 1 - t : t;  // alternate contexts walk the in-degree order in opposite directions
  const float4* emb4 = reinterpret_cast<const float4*>(a.atom_emb);
  uint32_t ph = 0, pph = 0;  // parities: per-step MMA barriers; plan buffers (bit b = buffer b)
  {  // the constant (1, 0, ..., 0) K-step that carries the biases: written once, never overwritten
    const uint32_t ones[8] = {0x00003c00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tc::tmem_st8(tOnes + lane_off, ones);
    tc::tmem_wait_st();
  }

  int buf = 0;
  for (int tile = ws.next_tile[0]; tile < n_tiles; tile = ws.next_tile[buf ^ 1], buf ^= 1) {
    if (t == 0) {  // next ticket; its record -> the other buffer (the readers of that buffer passed the end-of-tile barrier)
      const int nx = atomicAdd(a.tickets + tower, 1);
      ws.next_tile[buf ^ 1] = nx;  // read by every thread after the last barrier of this tile
      if (nx < n_tiles) {
        tc::fence_proxy_async_smem();
        tc::mbar_arrive_expect_tx(&ws.pbar[buf ^ 1], (uint32_t)sizeof(FusedTile));
        tc::bulk_copy_g2s(&ws.plan[buf ^ 1], tiles + nx, (uint32_t)sizeof(FusedTile), &ws.pbar[buf ^ 1]);
      }
    }
    F6_PROF(0);
    tc::mbar_wait(&ws.pbar[buf], (pph >> buf) & 1u);
    pph ^= 1u << buf;
    F6_PROF(1);
    const FusedTile& tp = ws.plan[buf];
    const uint32_t sw = tp.slot[myslot];
    const int r = sw & 127, deg = (sw >> 7) & 31, aid = (int)(sw >> 22);
    const uint32_t* entp = &tp.ent[(sw >> 12) & 1023];
    uint32_t* hbrow = &ws.hb[r * F6_HS];
    float h[D];
    {  // Embedding(atom): fp32 state in registers; packed once for the shared-memory copy (gathers) and the GEMM operand
      const float4* er = emb4 + aid * (D / 4);
      const float4* es = s_emb + aid * (D / 4);
      const int sw7 = aid & 7;
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 x = a.emb_smem ? es[c ^ sw7] : __ldg(er + c);
        h[4 * c] = x.x, h[4 * c + 1] = x.y, h[4 * c + 2] = x.z, h[4 * c + 3] = x.w;
        pk[2 * c] = tc::pack_f16x2(x.x, x.y), pk[2 * c + 1] = tc::pack_f16x2(x.z, x.w);
      }
#pragma unroll
      for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(hbrow)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      tc::tmem_st16(tAh + lane_off, pk);
      tc::tmem_wait_st();
    }
    tc::fence_before_thread_sync();
    F6_PROF(2);
    tc::named_bar_sync(bar_id, F6_THREADS);
    F6_PROF(3);

    for (int s = 0; s < a.steps; ++s) {
      const uint64_t dstep = (uint64_t)(s * (FusedPack6::BYTES / 16));
      const float* gb = reinterpret_cast<const float*>(smem + s * FusedPack6::BYTES + FusedPack6::OFF_BIAS);
      // ------------------------------------------------------------ Z in two K halves -> TMEM -> GEMM1 (-> gate GEMM)
#pragma unroll 1
      for (int hz = 0; hz < 2; ++hz) {
        __half2 acc[D * 2];
        // one entry: acc (+)= h[src][16 hz .. 16 hz + 16) (x) (mult * c[0..8)).  The K halves split the STATE columns, not the
        // bond components: every byte of a neighbour's row is gathered once per step (the gathers saturate the shared-memory
        // pipe during the Z phases: 16-byte reads of random rows conflict 2.15x)
        auto entry = [&](uint32_t ec, auto first_entry) {
          const uint4 cu = s_ctab[(ec >> 8) & 0xff];
          const uint32_t mbits = (ec >> 16) | (ec & 0xffff0000u);
          const __half2 mult = *reinterpret_cast<const __half2*>(&mbits);
          const __half2 c[4] = {__hmul2(*reinterpret_cast<const __half2*>(&cu.x), mult), __hmul2(*reinterpret_cast<const __half2*>(&cu.y), mult),
                                __hmul2(*reinterpret_cast<const __half2*>(&cu.z), mult), __hmul2(*reinterpret_cast<const __half2*>(&cu.w), mult)};
          const uint4* hp = reinterpret_cast<const uint4*>(&ws.hb[(ec & 0x7f) * F6_HS]) + 2 * hz;
#pragma unroll
          for (int q = 0; q < 2; ++q) {  // 8 columns per 16-byte read; HFMA2 broadcasts the low / high half
            const uint4 hv = hp[q];
            const __half2 hw[4] = {*reinterpret_cast<const __half2*>(&hv.x), *reinterpret_cast<const __half2*>(&hv.y),
                                   *reinterpret_cast<const __half2*>(&hv.z), *reinterpret_cast<const __half2*>(&hv.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __half2 lo = __low2half2(hw[i]), hi = __high2half2(hw[i]);
              const int m = 8 * q + 2 * i;  // column inside the half; accumulator (= TMEM column) 4 m + k / 2
#pragma unroll
              for (int kp = 0; kp < 4; ++kp) {
                if constexpr (decltype(first_entry)::value) {
                  acc[m * 4 + kp] = __hmul2(lo, c[kp]), acc[m * 4 + 4 + kp] = __hmul2(hi, c[kp]);
                } else {
                  acc[m * 4 + kp] = __hfma2(lo, c[kp], acc[m * 4 + kp]), acc[m * 4 + 4 + kp] = __hfma2(hi, c[kp], acc[m * 4 + 4 + kp]);
                }
              }
            }
          }
        };
        {
          // a row without entries runs the first-entry code on the all-zero descriptor (multiplicity 0: every product is 0):
          // a branch here is if-converted into 64 selects per half for EVERY warp (9 % of the kernel's instructions, ncu)
          const uint32_t e_first = deg > 0 ? entp[0] : 0u;
          uint32_t en = deg > 1 ? entp[1] : 0u;
          entry(e_first, F6True{});
#pragma unroll 1
          for (int e = 1; e < deg; ++e) {
            const uint32_t ec = en;
            if (e + 1 < deg) en = entp[e + 1];  // next entry's descriptor is in flight during this one's FMAs
            entry(ec, F6False{});
          }
        }
        F6_PROF(4);
        if (hz == 1) {  // GEMM1a must have consumed the first half before its columns are rewritten
          tc::mbar_wait(&ws.bar[3], ph);
          tc::fence_after_thread_sync();
        }
        F6_PROF(5);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = *reinterpret_cast<const uint32_t*>(&acc[ch * 32 + i]);
          tc::tmem_st32(tZ + lane_off + (uint32_t)(ch * 32), rr);
        }
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        F6_PROF(6);
        operands_ready();
        F6_PROF(7);
        if (mma_warp) {
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
        if constexpr (PRECISE || !F6_PACKED_GATES) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            rr[i] = tc::pack_f16x2(fz_sigmoid_half<PRECISE>(v[2 * i]) * h[2 * i], fz_sigmoid_half<PRECISE>(v[2 * i + 1]) * h[2 * i + 1]);
        } else {
          // packed: r * h = (0.5 tanh(y) + 0.5) h = fma(tanh(y), h/2, h/2) on the 16-bit copy of h (the product is rounded
          // to 16 bits for the operand anyway); one MUFU per TWO elements (tanh.approx.f16x2): the XU pipe issues one
          // warp instruction per 8 cycles and is the scarcest pipe of the gate phase
          const __half2 half = __floats2half2_rn(0.5f, 0.5f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 hv = reinterpret_cast<const uint4*>(hbrow)[c];
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __half2 hh = __hmul2(*reinterpret_cast<const __half2*>(&hw[i]), half);
              const __half2 th = fz_tanh_h2(__floats2half2_rn(v[8 * c + 2 * i], v[8 * c + 2 * i + 1]));
              const __half2 p = __hfma2(th, hh, hh);
              rr[4 * c + i] = *reinterpret_cast<const uint32_t*>(&p);
            }
          }
        }
        tc::tmem_st16(tAh + lane_off, rr);
      }
      tc::tmem_wait_st();
      tc::fence_before_thread_sync();
      F6_PROF(10);
      operands_ready();
      F6_PROF(11);
      // ------------------------------------------------------------ candidate GEMM: [r*h | 1] . [Wh_h ; bh] + agg . Wh_a
      if (mma_warp) {
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
        if constexpr (PRECISE || !F6_PACKED_GATES) {
#pragma unroll
          for (int j = 0; j < D; ++j) z[j] = fz_sigmoid_half<PRECISE>(v[j]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {  // one MUFU per two elements (tanh.approx.f16x2, 2^-11 absolute), then fp32
            const float2 tf = __half22float2(fz_tanh_h2(__floats2half2_rn(v[2 * i], v[2 * i + 1])));
            z[2 * i] = fmaf(0.5f, tf.x, 0.5f), z[2 * i + 1] = fmaf(0.5f, tf.y, 0.5f);
          }
        }
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
        const float inv = PRECISE ? 1.0f / sqrtf(var + a.eps) : fz_rsqrt_fast(var + a.eps);
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
      tc::named_bar_sync(bar_id, F6_THREADS);
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
        tc::named_bar_sync(bar_id, F6_THREADS);
        for (int mi = (t >> 4); mi < nm; mi += 8) {
          const int lo = tp.mol_lo[mi], hi = tp.mol_lo[mi + 1];
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four interleaved partial sums, combined in a fixed order
          int rr = lo;
          for (; rr + 4 <= hi; rr += 4) {
            s0 += hfp[rr * F6_HS + hl], s1 += hfp[(rr + 1) * F6_HS + hl];
            s2 += hfp[(rr + 2) * F6_HS + hl], s3 += hfp[(rr + 3) * F6_HS + hl];
          }
          for (; rr < hi; ++rr) s0 += hfp[rr * F6_HS + hl];
          a.pooled[(size_t)tp.molid[mi] * D + 16 * cb + hl] = (s0 + s1) + (s2 + s3);
        }
        tc::named_bar_sync(bar_id, F6_THREADS);
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
extern "C" void imp_debug_f6_prof(unsigned long long* host_out) {  // reads and clears the phase counters (profiling builds only)
  cudaMemcpyFromSymbol(host_out, imp::f6_prof_total, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {};
  cudaMemcpyToSymbol(imp::f6_prof_total, z, sizeof(z));
}
#endif

using namespace imp;

static int fused6_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

extern "C" int64_t imp_fused_pack_planned_bytes(int32_t d, int32_t bond_dim) {
  return (d == FZ_D && bond_dim == FZ_K) ? (int64_t)FusedPack6::BYTES : (int64_t)IMP_ERR_DIM;
}

static int fused_pack_planned_any(const float* d_bond_transform, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                                  void* d_packed, void* stream, int zperm7) {
  IMP_REQUIRE(d_bond_transform && d_packed && w && w->Wz && w->bz && w->Wr && w->br && w->Wh && w->bh && w->gamma && w->beta,
              IMP_ERR_ARG, "imp_fused_pack_planned: null pointer");
  IMP_REQUIRE(d == FZ_D && bond_dim == FZ_K, IMP_ERR_DIM, "imp_fused_pack_planned: built for atom_dim %d, bond_dim %d (got %d, %d)",
              FZ_D, FZ_K, d, bond_dim);
  const int n = FZ_D * FZ_D * FZ_K;
  fused_pack6_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_bond_transform, *w, (unsigned char*)d_packed, zperm7);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_fused_pack_planned(const float* d_bond_transform, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                                      void* d_packed, void* stream) {
  return fused_pack_planned_any(d_bond_transform, w, d, bond_dim, d_packed, stream, 0);
}

extern "C" int imp_fused_pack_planned7(const float* d_bond_transform, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                                       void* d_packed, void* stream) {
  return fused_pack_planned_any(d_bond_transform, w, d, bond_dim, d_packed, stream, 1);
}

namespace imp {
int launch_fused_h6(const void* d_plan, int32_t n_atoms, int32_t n_cat_atoms, int32_t bond_vocab, const float* d_atom_emb,
                    int32_t atom_vocab, const float* d_bond_emb, int32_t steps, const void* d_packed, float eps, bool precise,
                    float* d_pooled, cudaStream_t st) {
  Fused6Args a;
  a.plan = (const unsigned char*)d_plan, a.atom_emb = d_atom_emb, a.bond_emb = d_bond_emb, a.packed = (const unsigned char*)d_packed;
  a.pooled = d_pooled, a.atom_vocab = atom_vocab, a.bond_vocab = bond_vocab, a.steps = steps, a.eps = eps;
  // tile tickets: words 5, 6 of the plan header (FusedPlanHeader::pad), zeroed before every launch -- one forward at a time per plan
  a.tickets = reinterpret_cast<int*>(const_cast<void*>(d_plan)) + 5;
  IMP_CUDA(cudaMemsetAsync(a.tickets, 0, 2 * sizeof(int), st));
  // one persistent CTA per SM; CTAs are split between the towers in proportion to their atoms
  const int sms = fused6_sm_count();
  int nc = (int)((int64_t)sms * n_cat_atoms / (n_atoms > 0 ? n_atoms : 1));
  nc = nc < 1 ? 1 : (nc > sms - 1 ? sms - 1 : nc);
  int na = sms - nc;
  const int want = (int)ceil_div(ceil_div((int64_t)n_atoms, 100), F6_CTX) + 1;  // never more CTAs than a small batch has tiles for
  if (nc > want) nc = want;
  if (na > want) na = want;
  a.n_cta_cat = nc;
  size_t smem = (size_t)fused6_smem_bytes(steps, bond_vocab);
  IMP_REQUIRE(smem <= 227 * 1024, IMP_ERR_DIM, "imp_mpnn_forward_fused_planned: needs %zu B of shared memory", smem);
  a.emb_smem = smem + 128 + (size_t)atom_vocab * 128 <= 227 * 1024 ? 1 : 0;  // the atom-embedding table too, when it fits
  if (a.emb_smem) smem += 128 + (size_t)atom_vocab * 128;
  if (precise) {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h6_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h6_kernel<true><<<nc + na, F6_CTX * F6_THREADS, smem, st>>>(a);
  } else {
    IMP_CUDA(cudaFuncSetAttribute(mpnn_fused_h6_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mpnn_fused_h6_kernel<false><<<nc + na, F6_CTX * F6_THREADS, smem, st>>>(a);
  }
  IMP_LAUNCH_CHECK();
  return 0;
}
}  // namespace imp
