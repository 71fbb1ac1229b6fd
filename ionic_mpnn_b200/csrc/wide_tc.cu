// Wide atom states (atom_dim 256, bond_dim 8: BASELINE configs[4], the "wide/deep" variant) on the tensor cores.
//
// At d = 256 the forward is tensor-bound (238 FLOP/B), so the layers run as three pipelined tcgen05 GEMM
// kernels per message-passing step, all instances of ONE mainloop (wide_gemm_kernel<MODE>):
//
//   MSG   BondMatrixMessage o Reduce (models/layers.py:100-117,57-83):  agg = Z . Wc,  K = d * bond_dim = 2048, with
//         Z[v][m*8+k] = sum_e mult_e * c_e[k] * h[src_e][m]  (the exact re-association of DESIGN 4.1; c_e = bond embedding).
//         Z is never materialised: 256 producer threads (one per destination row) build each 256 x 64 slice of it straight
//         into the shared-memory A stage with packed HFMA2, while the Wc slice arrives by TMA bulk copy.
//   GRU1  z = sigma([h|agg] Wz + bz), r = sigma([h|agg] Wr + br)  (models/layers.py:144-148); the epilogue writes z and
//         r*h as 16-bit operands for GRU2.
//   GRU2  ht = tanh([r*h|agg] Wh + bh), n = (1-z) h + z ht, out = LayerNorm(n) gamma + beta + h  (layers.py:150-156):
//         the whole row (256 columns) sits in TMEM, so LayerNorm runs in the epilogue (n parked back into TMEM
//         between the statistics pass and the normalisation pass).
//
// Mainloop: a CTA owns a super-tile of 256 atom rows = two M = 128 accumulators (2 x 256 TMEM columns), so each
// 32 KB weight slice (N = 256, K = 64) feeds 8 tcgen05.mma of 128 x 256 x 16 -- the B bytes per flop of a 256 x 256
// tile, which is what keeps the L2 -> SM stream (the real ceiling of this shape: weights are 0.25-1 MB per layer and
// cannot be resident) at ~64 B/clk/SM.  3 stages x 64 KB of shared memory; warp 8 = TMA loader (cp.async.bulk +
// mbarrier byte counts), warp 9 = MMA issuer (elect.sync from warp-uniform code), warps 0-7 = producers (MSG) and
// epilogue (thread = atom row = TMEM lane).
//
// Activations live in "tile-packed" layouts that make every access of this pipeline contiguous: for a tile of 128
// rows, a 16-byte piece (8 halfs, or 4 floats) of all 128 rows is contiguous (2 KB), pieces follow each other.
//   TP16 (h16, agg16, z16, rh16): byte offset(row, piece p of 32) = (row >> 7) * 65536 + p * 2048 + (row & 127) * 16
//   TP32 (h32, the fp32 state):   byte offset(row, quad q of 64)  = (row >> 7) * 131072 + q * 2048 + (row & 127) * 16
// A K = 64 slice of a TP16 tile is 16 contiguous KB and already IS the canonical SWIZZLE_NONE K-major UMMA operand
// (LBO = 2048, SBO = 128): the A operand of the GRU GEMMs is one bulk copy per M tile.  Thread-per-row epilogue
// loads/stores are perfectly coalesced, and the gathers of the Z build (one 16-byte piece per entry per slice) fall
// into a 2 KB window per source tile.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace imp {
namespace wide {

constexpr int D = 256, KB = 8;
constexpr int TILE = 128, ST_ROWS = 256;
constexpr int KC = 64;                             // K per pipeline stage
constexpr int STAGES = 3;
constexpr int A_TILE_BYTES = TILE * KC * 2;        // 16 KB
constexpr int A_STAGE_BYTES = 2 * A_TILE_BYTES;    // two M tiles
constexpr int B_STAGE_BYTES = D * KC * 2;          // N = 256 rows x 64 k
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int THREADS = 320;                       // 8 worker warps + loader + MMA
constexpr int MSG_CHUNKS = D * KB / KC;            // 32
constexpr int GRU_CHUNKS = 2 * D / KC;             // 8
constexpr int MAX_BOND_VOCAB = 256;
constexpr int REG_ENTRIES = 4;                     // entries of a row decoded once per tile and kept in registers

// packed weights per (tower, step)
constexpr int64_t WC_BYTES = (int64_t)D * KB * D * 2;  // 1 MB
constexpr int64_t WG_BYTES = (int64_t)2 * D * D * 2;   // 256 KB per gate
constexpr int64_t OFF_WZ = WC_BYTES, OFF_WR = OFF_WZ + WG_BYTES, OFF_WH = OFF_WR + WG_BYTES, OFF_VEC = OFF_WH + WG_BYTES;
constexpr int64_t PACK_BYTES = OFF_VEC + 5 * D * 4;    // + bz, br, bh, gamma, beta (fp32)

enum { MODE_MSG = 0, MODE_GRU1 = 1, MODE_GRU2 = 2 };

struct Args {
  int n_atoms, n_cat;
  const int32_t *row_ptr, *col_src, *edge_bm;
  const float* bond_emb;
  int bond_vocab;
  const uint8_t* packed[2];  // this step's blob per tower
  uint8_t *h16, *agg16, *z16, *rh16;
  float* h32;
  float eps;
  int precise;
  long long* timeline;  // debug: phase boundaries of CTA 0 (tools/wide_timeline.py), normally null
};

struct Ctl {
  uint64_t full[STAGES], empty[STAGES], acc_full, acc_empty;
  uint32_t tmem;
};

__device__ __forceinline__ int64_t tp16_off(int row, int piece) { return (int64_t)(row >> 7) * 65536 + piece * 2048 + (row & 127) * 16; }
__device__ __forceinline__ int64_t tp32_off(int row, int quad) { return (int64_t)(row >> 7) * 131072 + quad * 2048 + (row & 127) * 16; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float fast_sigmoid(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float fast_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

// work item -> (tower, super-tile, gate half); the super-tile that straddles the tower boundary is visited once per tower
struct Item {
  int tower, st, half, lo, hi;
};
__device__ __forceinline__ Item decode_item(int i, int n_atoms, int n_cat, int halves) {
  Item it;
  it.half = i % halves;
  i /= halves;
  const int nc = (n_cat + ST_ROWS - 1) / ST_ROWS;
  if (i < nc) {
    it.tower = 0, it.st = i, it.lo = 0, it.hi = n_cat;
  } else {
    it.tower = 1, it.st = n_cat / ST_ROWS + (i - nc), it.lo = n_cat, it.hi = n_atoms;
  }
  return it;
}
static inline int n_items(int n_atoms, int n_cat) {
  const int nc = (n_cat + ST_ROWS - 1) / ST_ROWS;
  const int na = n_atoms > n_cat ? (n_atoms + ST_ROWS - 1) / ST_ROWS - n_cat / ST_ROWS : 0;
  return nc + na;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) wide_gemm_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [STAGES x (A tile 0 | A tile 1 | B)] [ctab: bond_vocab x 8 halfs] [Ctl]
  uint8_t* stage_base = smem;
  __half* ctab = reinterpret_cast<__half*>(smem + STAGES * STAGE_BYTES);
  Ctl& ctl = *reinterpret_cast<Ctl*>(smem + STAGES * STAGE_BYTES + MAX_BOND_VOCAB * 16);
  constexpr int HALVES = MODE == MODE_GRU1 ? 2 : 1;
  constexpr int CHUNKS = MODE == MODE_MSG ? MSG_CHUNKS : GRU_CHUNKS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = (((a.n_cat + ST_ROWS - 1) / ST_ROWS) +
                     (a.n_atoms > a.n_cat ? (a.n_atoms + ST_ROWS - 1) / ST_ROWS - a.n_cat / ST_ROWS : 0)) * HALVES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&ctl.full[s], MODE == MODE_MSG ? 9 : 1);
      tc::mbar_init(&ctl.empty[s], 1);
    }
    tc::mbar_init(&ctl.acc_full, 1);
    tc::mbar_init(&ctl.acc_empty, 8);
    tc::mbar_fence_init();
  }
  if (MODE == MODE_MSG) {
    for (int i = threadIdx.x; i < a.bond_vocab * KB; i += THREADS) ctab[i] = __float2half_rn(a.bond_emb[i]);
  }
  if (warp == 9) tc::tmem_alloc<512>(&ctl.tmem);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = ctl.tmem;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA loader
    if (lane == 0) {
      uint32_t it = 0;
      for (int i = blockIdx.x; i < items; i += gridDim.x) {
        const Item w = decode_item(i, a.n_atoms, a.n_cat, HALVES);
        const uint8_t* bsrc = (w.tower ? a.packed[1] : a.packed[0]) + (MODE == MODE_MSG ? 0 : MODE == MODE_GRU1 ? (w.half ? OFF_WR : OFF_WZ) : OFF_WH);
        const uint8_t* x0 = MODE == MODE_GRU1 ? a.h16 : a.rh16;
        for (int j = 0; j < CHUNKS; ++j, ++it) {
          const int s = it % STAGES;
          tc::mbar_wait(&ctl.empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* sb = stage_base + s * STAGE_BYTES;
          tc::mbar_arrive_expect_tx(&ctl.full[s], MODE == MODE_MSG ? B_STAGE_BYTES : STAGE_BYTES);
          tc::bulk_copy_g2s(sb + A_STAGE_BYTES, bsrc + (int64_t)j * B_STAGE_BYTES, B_STAGE_BYTES, &ctl.full[s]);
          if (MODE != MODE_MSG) {
            const uint8_t* x = j < GRU_CHUNKS / 2 ? x0 : a.agg16;
            const int jj = j < GRU_CHUNKS / 2 ? j : j - GRU_CHUNKS / 2;
            const uint8_t* src = x + (int64_t)(w.st * 2) * 65536 + (int64_t)jj * A_TILE_BYTES;
            tc::bulk_copy_g2s(sb, src, A_TILE_BYTES, &ctl.full[s]);
            tc::bulk_copy_g2s(sb + A_TILE_BYTES, src + 65536, A_TILE_BYTES, &ctl.full[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::make_idesc(tc::FMT_F16, TILE, D);
    uint32_t it = 0, k = 0;
    for (int i = blockIdx.x; i < items; i += gridDim.x, ++k) {
      tc::mbar_wait(&ctl.acc_empty, (k & 1) ^ 1);
      tc::fence_after_thread_sync();
      for (int j = 0; j < CHUNKS; ++j, ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(&ctl.full[s], (it / STAGES) & 1);
        tc::fence_after_thread_sync();
        const uint32_t sa = tc::smem_u32(stage_base + s * STAGE_BYTES);
        const uint64_t da = tc::make_smem_desc(sa, 2048, 128), db = tc::make_smem_desc(sa + A_STAGE_BYTES, 4096, 128);
        if (tc::elect_one()) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              tc::mma_bf16(tmem + mt * D, da + (uint64_t)((mt * A_TILE_BYTES + ks * 4096) >> 4), db + (uint64_t)((ks * 8192) >> 4),
                           idesc, j > 0 || ks > 0);
          tc::mma_commit(&ctl.empty[s]);
          if (j == CHUNKS - 1) tc::mma_commit(&ctl.acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ workers: Z build (MSG) + epilogue
    const int tid = threadIdx.x;  // 0..255 = row within the super-tile
    const int mt = tid >> 7;
    const uint32_t tacc = tmem + mt * D + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t it = 0, k = 0;
    for (int i = blockIdx.x; i < items; i += gridDim.x, ++k) {
      const Item w = decode_item(i, a.n_atoms, a.n_cat, HALVES);
      const int row = w.st * ST_ROWS + tid;
      const bool mine = row >= w.lo && row < w.hi;
      const float* vec = reinterpret_cast<const float*>((w.tower ? a.packed[1] : a.packed[0]) + OFF_VEC);
      if (MODE == MODE_MSG) {
        int e0 = 0, e1 = 0;
        if (mine) e0 = __ldg(a.row_ptr + row), e1 = __ldg(a.row_ptr + row + 1);
        const int deg = e1 - e0;
        int64_t soff[REG_ENTRIES];
        __half2 cm[REG_ENTRIES][4];
#pragma unroll
        for (int e = 0; e < REG_ENTRIES; ++e) {
          soff[e] = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) cm[e][q] = __float2half2_rn(0.f);
          if (e < deg) {
            const int src = __ldg(a.col_src + e0 + e);
            const uint32_t bm = (uint32_t)__ldg(a.edge_bm + e0 + e);
            soff[e] = tp16_off(src, 0);
            const __half2 mult = __float2half2_rn((float)(bm >> 16));
            const uint4 cw = *reinterpret_cast<const uint4*>(ctab + (bm & 0xffffu) * KB);
            const __half2* c2 = reinterpret_cast<const __half2*>(&cw);
#pragma unroll
            for (int q = 0; q < 4; ++q) cm[e][q] = __hmul2(c2[q], mult);
          }
        }
        // the gathers of slice j + 1 are in flight while slice j is accumulated (they do not depend on a stage being free)
        uint4 hw[REG_ENTRIES], hnext[REG_ENTRIES];
#pragma unroll
        for (int e = 0; e < REG_ENTRIES; ++e) {
          hnext[e] = make_uint4(0, 0, 0, 0);
          if (e < deg) hnext[e] = __ldg(reinterpret_cast<const uint4*>(a.h16 + soff[e]));
        }
        for (int j = 0; j < MSG_CHUNKS; ++j, ++it) {
          const int s = it % STAGES;
          __half2 acc[8][4];
#pragma unroll
          for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[p][q] = __float2half2_rn(0.f);
#pragma unroll
          for (int e = 0; e < REG_ENTRIES; ++e) {
            hw[e] = hnext[e];
            if (e < deg && j + 1 < MSG_CHUNKS) hnext[e] = __ldg(reinterpret_cast<const uint4*>(a.h16 + soff[e] + (j + 1) * 2048));
          }
#pragma unroll
          for (int e = 0; e < REG_ENTRIES; ++e) {
            if (e < deg) {
              const __half2* h2 = reinterpret_cast<const __half2*>(&hw[e]);
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                const __half2 lo = __low2half2(h2[p]), hi = __high2half2(h2[p]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  acc[2 * p][q] = __hfma2(lo, cm[e][q], acc[2 * p][q]);
                  acc[2 * p + 1][q] = __hfma2(hi, cm[e][q], acc[2 * p + 1][q]);
                }
              }
            }
          }
          for (int e = e0 + REG_ENTRIES; e < e1; ++e) {  // rows with more than REG_ENTRIES unique neighbours
            const int src = __ldg(a.col_src + e);
            const uint32_t bm = (uint32_t)__ldg(a.edge_bm + e);
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(a.h16 + tp16_off(src, j)));
            const __half2 mult = __float2half2_rn((float)(bm >> 16));
            const uint4 cw = *reinterpret_cast<const uint4*>(ctab + (bm & 0xffffu) * KB);
            const __half2* c2 = reinterpret_cast<const __half2*>(&cw);
            const __half2* h2 = reinterpret_cast<const __half2*>(&x);
            __half2 c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = __hmul2(c2[q], mult);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const __half2 lo = __low2half2(h2[p]), hi = __high2half2(h2[p]);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                acc[2 * p][q] = __hfma2(lo, c[q], acc[2 * p][q]);
                acc[2 * p + 1][q] = __hfma2(hi, c[q], acc[2 * p + 1][q]);
              }
            }
          }
          tc::mbar_wait(&ctl.empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* dst = stage_base + s * STAGE_BYTES + mt * A_TILE_BYTES + (tid & 127) * 16;
#pragma unroll
          for (int p = 0; p < 8; ++p) *reinterpret_cast<uint4*>(dst + p * 2048) = *reinterpret_cast<const uint4*>(&acc[p][0]);
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl.full[s]);
        }
      }
      // ---- epilogue
      tc::mbar_wait(&ctl.acc_full, k & 1);
      tc::fence_after_thread_sync();
      if (MODE == MODE_MSG) {
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float v[32];
          tc::tmem_ld32(tacc + c * 32, v);
          if (mine) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
              *reinterpret_cast<uint4*>(a.agg16 + tp16_off(row, 4 * c + p)) =
                  make_uint4(tc::pack_f16x2(v[8 * p], v[8 * p + 1]), tc::pack_f16x2(v[8 * p + 2], v[8 * p + 3]),
                             tc::pack_f16x2(v[8 * p + 4], v[8 * p + 5]), tc::pack_f16x2(v[8 * p + 6], v[8 * p + 7]));
          }
        }
      } else if (MODE == MODE_GRU1) {
        const float* bias = vec + (w.half ? D : 0);
        uint8_t* out = w.half ? a.rh16 : a.z16;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float v[32];
          tc::tmem_ld32(tacc + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = v[i] + __ldg(bias + c * 32 + i);
            v[i] = a.precise ? 1.0f / (1.0f + expf(-x)) : fast_sigmoid(x);
          }
          if (mine) {
            if (w.half) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 h = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(a.h32) + tp32_off(row, 8 * c + q));
                v[4 * q] *= h.x, v[4 * q + 1] *= h.y, v[4 * q + 2] *= h.z, v[4 * q + 3] *= h.w;
              }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
              *reinterpret_cast<uint4*>(out + tp16_off(row, 4 * c + p)) =
                  make_uint4(tc::pack_f16x2(v[8 * p], v[8 * p + 1]), tc::pack_f16x2(v[8 * p + 2], v[8 * p + 3]),
                             tc::pack_f16x2(v[8 * p + 4], v[8 * p + 5]), tc::pack_f16x2(v[8 * p + 6], v[8 * p + 7]));
          }
        }
      } else {
        const float *bh = vec + 2 * D, *gamma = vec + 3 * D, *beta = vec + 4 * D;
        uint8_t* h32b = reinterpret_cast<uint8_t*>(a.h32);
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float v[32];
          tc::tmem_ld32(tacc + c * 32, v);
          uint32_t nn[32];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            uint4 zw = make_uint4(0, 0, 0, 0);
            if (mine) zw = *reinterpret_cast<const uint4*>(a.z16 + tp16_off(row, 4 * c + p));
            const __half2* z2 = reinterpret_cast<const __half2*>(&zw);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
              if (mine) h = *reinterpret_cast<const float4*>(h32b + tp32_off(row, 8 * c + 2 * p + q));
              const float2 za = __half22float2(z2[2 * q]), zb = __half22float2(z2[2 * q + 1]);
              const float hh[4] = {h.x, h.y, h.z, h.w}, zz[4] = {za.x, za.y, zb.x, zb.y};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int i = 8 * p + 4 * q + t;
                const float x = v[i] + __ldg(bh + c * 32 + i);
                const float ht = a.precise ? tanhf(x) : fast_tanh(x);
                const float n = fmaf(zz[t], ht - hh[t], hh[t]);
                sum += n, sq = fmaf(n, n, sq);
                nn[i] = __float_as_uint(n);
              }
            }
          }
          tc::tmem_st32(tacc + c * 32, nn);
        }
        tc::tmem_wait_st();
        const float mean = sum * (1.0f / D);
        const float var = fmaxf(sq * (1.0f / D) - mean * mean, 0.f);
        const float rstd = a.precise ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float v[32];
          tc::tmem_ld32(tacc + c * 32, v);
          if (mine) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4* hp = reinterpret_cast<float4*>(h32b + tp32_off(row, 8 * c + q));
              const float4 h = *hp;
              const int i = 4 * q;
              float4 y;
              y.x = fmaf((v[i] - mean) * rstd, __ldg(gamma + c * 32 + i), __ldg(beta + c * 32 + i)) + h.x;
              y.y = fmaf((v[i + 1] - mean) * rstd, __ldg(gamma + c * 32 + i + 1), __ldg(beta + c * 32 + i + 1)) + h.y;
              y.z = fmaf((v[i + 2] - mean) * rstd, __ldg(gamma + c * 32 + i + 2), __ldg(beta + c * 32 + i + 2)) + h.z;
              y.w = fmaf((v[i + 3] - mean) * rstd, __ldg(gamma + c * 32 + i + 3), __ldg(beta + c * 32 + i + 3)) + h.w;
              *hp = y;
              v[i] = y.x, v[i + 1] = y.y, v[i + 2] = y.z, v[i + 3] = y.w;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
              *reinterpret_cast<uint4*>(a.h16 + tp16_off(row, 4 * c + p)) =
                  make_uint4(tc::pack_f16x2(v[8 * p], v[8 * p + 1]), tc::pack_f16x2(v[8 * p + 2], v[8 * p + 3]),
                             tc::pack_f16x2(v[8 * p + 4], v[8 * p + 5]), tc::pack_f16x2(v[8 * p + 6], v[8 * p + 7]));
          }
        }
      }
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.acc_empty);
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 9) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GatedUpdate.call (models/layers.py:142-156) as ONE kernel per step: a CTA owns a 128-row tile whose gates never leave the
// SM.  TMEM: z pre-activations in columns [0,256) (later the row's fp32 state), r in [256,512) (later the candidate,
// later the blended row n).  Shared memory: a resident 64 KB operand tile X (TP16 layout = UMMA layout) + 3 stages.
//   tile start  X <- the tile's h16 (one 64 KB bulk copy); the tile's fp32 state (128 contiguous KB) is prefetched into L2
//   phase A     [h|agg] . [Wz|Wr]: 16 K-slices of 32; A from X (K < 256) or a streamed agg slice, both weight slices streamed
//   EA          r = sigma(. + br); X <- r * h in place (X is now the A operand of phase B)
//   phase B     [r*h|agg] . Wh -> the r columns
//   E2 pass 1   z = sigma(. + bz), ht = tanh(. + bh), n = h + z (ht - h); n parked over the candidate and h (fp32, read once
//               from L2) over the z columns; LayerNorm partial sums (two workers per row meet through shared memory)
//   E2 pass 2   out = (n - mean) rstd gamma + beta + h from TMEM only; fp32 state and its 16-bit copy written in place.
// Every epilogue loop is rolled (32 / 16 columns per trip): an unrolled 128-column body per thread is ~6,000 SASS
// instructions and ran instruction-fetch bound (ncu: 15 no_instruction stalls per issue, profiles/README.md).
constexpr int G_KC = 32, G_STAGES = 3, G_CHUNKS = 2 * D / G_KC;  // 16 slices per phase
constexpr int G_A_BYTES = TILE * G_KC * 2;                        // 8 KB
constexpr int G_B_BYTES = D * G_KC * 2;                           // 16 KB per gate
constexpr int G_STAGE_BYTES = G_A_BYTES + 2 * G_B_BYTES;          // 40 KB
constexpr int G_X_BYTES = TILE * D * 2;                           // 64 KB
constexpr int G_CH = 4, G_WORKERS = 4 * G_CH, G_COLS = D / G_CH;  // worker warps: 4 lane quarters x G_CH column slices
constexpr int G_THREADS = (G_WORKERS + 2) * 32;

struct GCtl {
  uint64_t full[G_STAGES], empty[G_STAGES], x_full, accA, rh_ready, accB, acc_empty;
  uint32_t tmem;
};
constexpr int G_OFF_X = G_STAGES * G_STAGE_BYTES, G_OFF_VEC = G_OFF_X + G_X_BYTES, G_OFF_RED = G_OFF_VEC + 5 * D * 4,
              G_OFF_CTL = G_OFF_RED + G_CH * TILE * 8;
constexpr int G_SMEM_BYTES = G_OFF_CTL + (int)sizeof(GCtl) + 64;

struct GItem {
  int tower, t, lo, hi;
};
// Work items = (tower, 128-row tile).  CL == 2: each tower's list is padded to an even length, so that the two CTAs of a
// cluster (items 2m, 2m + 1) always work on the same tower's weights; a padding item re-reads the tower's last tile and
// owns no row.
template <int CL>
__host__ __device__ __forceinline__ int n_gitems(int n_atoms, int n_cat) {
  const int nc = (n_cat + TILE - 1) / TILE;
  const int na = n_atoms > n_cat ? (n_atoms + TILE - 1) / TILE - n_cat / TILE : 0;
  return CL == 2 ? ((nc + 1) & ~1) + ((na + 1) & ~1) : nc + na;
}
template <int CL>
__device__ __forceinline__ GItem decode_gitem(int i, int n_atoms, int n_cat) {
  GItem it;
  const int nc = (n_cat + TILE - 1) / TILE;
  const int na = n_atoms > n_cat ? (n_atoms + TILE - 1) / TILE - n_cat / TILE : 0;
  const int np0 = CL == 2 ? (nc + 1) & ~1 : nc;
  if (i < np0) {
    const bool valid = i < nc;
    it.tower = 0, it.t = valid ? i : nc - 1, it.lo = 0, it.hi = valid ? n_cat : 0;
  } else {
    const int j = i - np0;
    const bool valid = j < na;
    it.tower = 1, it.t = n_cat / TILE + (valid ? j : na - 1), it.lo = n_cat, it.hi = valid ? n_atoms : n_cat;
  }
  return it;
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

template <bool PRECISE>
__device__ __forceinline__ float sigmoid_f(float x) {
  return PRECISE ? 1.0f / (1.0f + expf(-x)) : fast_sigmoid(x);
}
template <bool PRECISE>
__device__ __forceinline__ float tanh_f(float x) {
  return PRECISE ? tanhf(x) : fast_tanh(x);
}

template <bool PRECISE, int CL>
__global__ void __launch_bounds__(G_THREADS, 1) wide_gru_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* xs = smem + G_OFF_X;
  float* vec_s = reinterpret_cast<float*>(smem + G_OFF_VEC);
  float2* red_s = reinterpret_cast<float2*>(smem + G_OFF_RED);
  GCtl& ctl = *reinterpret_cast<GCtl*>(smem + G_OFF_CTL);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = n_gitems<CL>(a.n_atoms, a.n_cat);
  const uint32_t rank = CL == 2 ? tc::cluster_ctarank() : 0;  // CL == 2: a pair of CTAs (two tiles of one tower) shares every
                                                               // weight slice: each loads half and multicasts it to both

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      tc::mbar_init(&ctl.full[s], 1);
      tc::mbar_init(&ctl.empty[s], CL);  // a stage is free when every CTA that receives multicast data in it has consumed it
    }
    tc::mbar_init(&ctl.x_full, 1);
    tc::mbar_init(&ctl.accA, 1);
    tc::mbar_init(&ctl.rh_ready, G_WORKERS);
    tc::mbar_init(&ctl.accB, 1);
    tc::mbar_init(&ctl.acc_empty, G_WORKERS);
    tc::mbar_fence_init();
  }
  if (warp == G_WORKERS + 1) tc::tmem_alloc<512>(&ctl.tmem);
  tc::fence_before_thread_sync();
  __syncthreads();
  if (CL == 2) tc::cluster_sync();  // the peer's barriers exist before any multicast copy or commit can reach them
  tc::fence_after_thread_sync();
  const uint32_t tmem = ctl.tmem;

  if (warp == G_WORKERS) {
    // ------------------------------------------------------------------ TMA loader
    if (lane == 0) {
      uint32_t it = 0, k = 0;
      for (int i = blockIdx.x; i < items; i += gridDim.x, ++k) {
        const GItem w = decode_gitem<CL>(i, a.n_atoms, a.n_cat);
        const uint8_t* pk = w.tower ? a.packed[1] : a.packed[0];
        const uint8_t* aggt = a.agg16 + (int64_t)w.t * 65536;
        if (k > 0) tc::mbar_wait(&ctl.accB, (k - 1) & 1);  // phase B of the previous tile has finished reading X
        tc::mbar_arrive_expect_tx(&ctl.x_full, G_X_BYTES);
        tc::bulk_copy_g2s(xs, a.h16 + (int64_t)w.t * 65536, G_X_BYTES, &ctl.x_full);
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(a.h32) + (int64_t)w.t * 131072, 131072);
        for (int j = 0; j < 2 * G_CHUNKS; ++j, ++it) {  // phase A: slices 0..15, phase B: 16..31
          const int s = it % G_STAGES;
          tc::mbar_wait(&ctl.empty[s], ((it / G_STAGES) & 1) ^ 1);
          uint8_t* sb = smem + s * G_STAGE_BYTES;
          const bool phase_b = j >= G_CHUNKS;
          const int jj = phase_b ? j - G_CHUNKS : j;
          const bool stream_a = jj >= G_CHUNKS / 2;
          tc::mbar_arrive_expect_tx(&ctl.full[s], (phase_b ? G_B_BYTES : 2 * G_B_BYTES) + (stream_a ? G_A_BYTES : 0));
          if (stream_a) tc::bulk_copy_g2s(sb, aggt + (jj - G_CHUNKS / 2) * G_A_BYTES, G_A_BYTES, &ctl.full[s]);
          if (CL == 2) {  // this CTA's half of the weight slice(s), into the same stage of both CTAs
            if (phase_b)
              tc::bulk_copy_g2s_mc(sb + G_A_BYTES + rank * (G_B_BYTES / 2), pk + OFF_WH + (int64_t)jj * G_B_BYTES + rank * (G_B_BYTES / 2),
                                   G_B_BYTES / 2, &ctl.full[s], 3);
            else
              tc::bulk_copy_g2s_mc(sb + G_A_BYTES + rank * G_B_BYTES, pk + (rank ? OFF_WR : OFF_WZ) + (int64_t)jj * G_B_BYTES, G_B_BYTES,
                                   &ctl.full[s], 3);
          } else if (phase_b) {
            tc::bulk_copy_g2s(sb + G_A_BYTES, pk + OFF_WH + (int64_t)jj * G_B_BYTES, G_B_BYTES, &ctl.full[s]);
          } else {
            tc::bulk_copy_g2s(sb + G_A_BYTES, pk + OFF_WZ + (int64_t)jj * G_B_BYTES, G_B_BYTES, &ctl.full[s]);
            tc::bulk_copy_g2s(sb + G_A_BYTES + G_B_BYTES, pk + OFF_WR + (int64_t)jj * G_B_BYTES, G_B_BYTES, &ctl.full[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == G_WORKERS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::make_idesc(tc::FMT_F16, TILE, D);
    const uint32_t x_addr = tc::smem_u32(xs);
    uint32_t it = 0, k = 0;
    for (int i = blockIdx.x; i < items; i += gridDim.x, ++k) {
      tc::mbar_wait(&ctl.acc_empty, (k & 1) ^ 1);
      tc::mbar_wait(&ctl.x_full, k & 1);
      tc::fence_after_thread_sync();
      for (int j = 0; j < G_CHUNKS; ++j, ++it) {
        const int s = it % G_STAGES;
        tc::mbar_wait(&ctl.full[s], (it / G_STAGES) & 1);
        tc::fence_after_thread_sync();
        const uint32_t sa = tc::smem_u32(smem + s * G_STAGE_BYTES);
        const uint64_t da = tc::make_smem_desc(j < G_CHUNKS / 2 ? x_addr + j * G_A_BYTES : sa, 2048, 128),
                       dz = tc::make_smem_desc(sa + G_A_BYTES, 4096, 128), dr = tc::make_smem_desc(sa + G_A_BYTES + G_B_BYTES, 4096, 128);
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < G_KC / 16; ++ks) {
            tc::mma_bf16(tmem, da + (uint64_t)((ks * 4096) >> 4), dz + (uint64_t)((ks * 8192) >> 4), idesc, j > 0 || ks > 0);
            tc::mma_bf16(tmem + D, da + (uint64_t)((ks * 4096) >> 4), dr + (uint64_t)((ks * 8192) >> 4), idesc, j > 0 || ks > 0);
          }
          if (CL == 2) tc::mma_commit_mc(&ctl.empty[s], 3); else tc::mma_commit(&ctl.empty[s]);
          if (j == G_CHUNKS - 1) tc::mma_commit(&ctl.accA);
        }
        __syncwarp();
      }
      tc::mbar_wait(&ctl.rh_ready, k & 1);
      tc::fence_after_thread_sync();
      for (int j = 0; j < G_CHUNKS; ++j, ++it) {
        const int s = it % G_STAGES;
        tc::mbar_wait(&ctl.full[s], (it / G_STAGES) & 1);
        tc::fence_after_thread_sync();
        const uint32_t sa = tc::smem_u32(smem + s * G_STAGE_BYTES);
        const uint64_t da = tc::make_smem_desc(j < G_CHUNKS / 2 ? x_addr + j * G_A_BYTES : sa, 2048, 128),
                       db = tc::make_smem_desc(sa + G_A_BYTES, 4096, 128);
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < G_KC / 16; ++ks)
            tc::mma_bf16(tmem + D, da + (uint64_t)((ks * 4096) >> 4), db + (uint64_t)((ks * 8192) >> 4), idesc, j > 0 || ks > 0);
          if (CL == 2) tc::mma_commit_mc(&ctl.empty[s], 3); else tc::mma_commit(&ctl.empty[s]);
          if (j == G_CHUNKS - 1) tc::mma_commit(&ctl.accB);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ workers: 4 lane quarters x G_CH column slices
    const int q = warp & 3, ch = warp >> 2;
    const int rowt = q * 32 + lane;
    const uint32_t tz = tmem + (uint32_t)(ch * G_COLS) + ((uint32_t)(q * 32) << 16), tr = tz + D;
    const float *bz = vec_s + ch * G_COLS, *br = vec_s + D + ch * G_COLS, *bh = vec_s + 2 * D + ch * G_COLS,
                *gamma = vec_s + 3 * D + ch * G_COLS, *beta = vec_s + 4 * D + ch * G_COLS;
    uint8_t* h32b = reinterpret_cast<uint8_t*>(a.h32);
    uint8_t* xrow = xs + (G_COLS / 8 * ch) * 2048 + rowt * 16;  // this thread's pieces of X: + p * 2048
    int cur_tower = -1;
    uint32_t k = 0;
    for (int i = blockIdx.x; i < items; i += gridDim.x, ++k) {
      const GItem w = decode_gitem<CL>(i, a.n_atoms, a.n_cat);
      const int row = w.t * TILE + rowt;
      const bool mine = row >= w.lo && row < w.hi;
      if (w.tower != cur_tower) {  // bz, br, bh, gamma, beta of this tower -> shared memory
        tc::named_bar_sync(1, G_WORKERS * 32);
        const float* v = reinterpret_cast<const float*>((w.tower ? a.packed[1] : a.packed[0]) + OFF_VEC);
        for (int x = threadIdx.x; x < 5 * D; x += G_WORKERS * 32) vec_s[x] = __ldg(v + x);
        tc::named_bar_sync(1, G_WORKERS * 32);
        cur_tower = w.tower;
      }
      const uint8_t* hrow = h32b + tp32_off(row, G_COLS / 4 * ch);  // this thread's quads of the fp32 state: + x * 2048
      // ---- EA: r -> r*h operand, in place over the h16 tile
      const bool tl = a.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && k < 16;
      if (tl) a.timeline[k * 8 + 0] = clock64();
      tc::mbar_wait(&ctl.x_full, k & 1);
      tc::mbar_wait(&ctl.accA, k & 1);
      tc::fence_after_thread_sync();
      if (tl) a.timeline[k * 8 + 1] = clock64();
#pragma unroll 1
      for (int c = 0; c < G_COLS / 32; ++c) {
        float v[32];
        tc::tmem_ld32(tr + c * 32, v);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          uint4* xp = reinterpret_cast<uint4*>(xrow + (4 * c + p) * 2048);
          const uint4 hw = *xp;
          const __half2* h2 = reinterpret_cast<const __half2*>(&hw);
          const float4 b0 = *reinterpret_cast<const float4*>(br + c * 32 + 8 * p), b1 = *reinterpret_cast<const float4*>(br + c * 32 + 8 * p + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float2 hf = __half22float2(h2[x]);
            o[2 * x] = sigmoid_f<PRECISE>(v[8 * p + 2 * x] + bb[2 * x]) * hf.x;
            o[2 * x + 1] = sigmoid_f<PRECISE>(v[8 * p + 2 * x + 1] + bb[2 * x + 1]) * hf.y;
          }
          *xp = make_uint4(tc::pack_f16x2(o[0], o[1]), tc::pack_f16x2(o[2], o[3]), tc::pack_f16x2(o[4], o[5]), tc::pack_f16x2(o[6], o[7]));
        }
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.rh_ready);
      if (tl) a.timeline[k * 8 + 2] = clock64();
      // ---- E2 pass 1: blend, statistics; n -> candidate columns, h -> z columns
      float4 hq[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) hq[x] = mine ? *reinterpret_cast<const float4*>(hrow + x * 2048) : make_float4(0.f, 0.f, 0.f, 0.f);
      tc::mbar_wait(&ctl.accB, k & 1);
      tc::fence_after_thread_sync();
      if (tl) a.timeline[k * 8 + 3] = clock64();
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < G_COLS / 16; ++c) {
        float zp[16], hp[16];
        tc::tmem_ld16(tz + c * 16, zp);
        tc::tmem_ld16(tr + c * 16, hp);
        float4 hn[4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
          hn[x] = (mine && c < G_COLS / 16 - 1) ? *reinterpret_cast<const float4*>(hrow + (4 * (c + 1) + x) * 2048) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t nn[16], hh[16];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float4 b0 = *reinterpret_cast<const float4*>(bz + c * 16 + 4 * x), b1 = *reinterpret_cast<const float4*>(bh + c * 16 + 4 * x);
          const float bzv[4] = {b0.x, b0.y, b0.z, b0.w}, bhv[4] = {b1.x, b1.y, b1.z, b1.w}, hv[4] = {hq[x].x, hq[x].y, hq[x].z, hq[x].w};
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            const float z = sigmoid_f<PRECISE>(zp[4 * x + y] + bzv[y]), ht = tanh_f<PRECISE>(hp[4 * x + y] + bhv[y]);
            const float n = fmaf(z, ht - hv[y], hv[y]);
            sum += n, sq = fmaf(n, n, sq);
            nn[4 * x + y] = __float_as_uint(n), hh[4 * x + y] = __float_as_uint(hv[y]);
          }
        }
        tc::tmem_st16(tr + c * 16, nn);
        tc::tmem_st16(tz + c * 16, hh);
#pragma unroll
        for (int x = 0; x < 4; ++x) hq[x] = hn[x];
      }
      red_s[ch * TILE + rowt] = make_float2(sum, sq);
      tc::tmem_wait_st();
      tc::named_bar_sync(1, G_WORKERS * 32);
      if (tl) a.timeline[k * 8 + 4] = clock64();
      sum = 0.f, sq = 0.f;
#pragma unroll
      for (int o = 0; o < G_CH; ++o) {  // the same order in every slice: all workers of a row see identical statistics
        const float2 t2 = red_s[o * TILE + rowt];
        sum += t2.x, sq += t2.y;
      }
      const float mean = sum * (1.0f / D);
      const float var = fmaxf(sq * (1.0f / D) - mean * mean, 0.f);
      const float rstd = PRECISE ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
      // ---- E2 pass 2: normalise + residual from TMEM, write the state
#pragma unroll 1
      for (int c = 0; c < G_COLS / 16; ++c) {
        float v[16], hv[16];
        tc::tmem_ld16(tr + c * 16, v);
        tc::tmem_ld16(tz + c * 16, hv);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float4 g = *reinterpret_cast<const float4*>(gamma + c * 16 + 4 * x), b = *reinterpret_cast<const float4*>(beta + c * 16 + 4 * x);
          v[4 * x] = fmaf((v[4 * x] - mean) * rstd, g.x, b.x) + hv[4 * x];
          v[4 * x + 1] = fmaf((v[4 * x + 1] - mean) * rstd, g.y, b.y) + hv[4 * x + 1];
          v[4 * x + 2] = fmaf((v[4 * x + 2] - mean) * rstd, g.z, b.z) + hv[4 * x + 2];
          v[4 * x + 3] = fmaf((v[4 * x + 3] - mean) * rstd, g.w, b.w) + hv[4 * x + 3];
        }
        if (mine) {
#pragma unroll
          for (int x = 0; x < 4; ++x)
            *reinterpret_cast<float4*>(h32b + tp32_off(row, G_COLS / 4 * ch + 4 * c + x)) = make_float4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
#pragma unroll
          for (int p = 0; p < 2; ++p)
            *reinterpret_cast<uint4*>(a.h16 + tp16_off(row, G_COLS / 8 * ch + 2 * c + p)) =
                make_uint4(tc::pack_f16x2(v[8 * p], v[8 * p + 1]), tc::pack_f16x2(v[8 * p + 2], v[8 * p + 3]),
                           tc::pack_f16x2(v[8 * p + 4], v[8 * p + 5]), tc::pack_f16x2(v[8 * p + 6], v[8 * p + 7]));
        }
      }
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.acc_empty);
      if (tl) a.timeline[k * 8 + 5] = clock64();
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (CL == 2) tc::cluster_sync();  // no CTA leaves while its peer can still signal its barriers
  if (warp == G_WORKERS + 1) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc<512>(tmem);
  }
}

constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + MAX_BOND_VOCAB * 16 + (int)sizeof(Ctl) + 64;

// Embedding(atom) (train_viscosity.py:163,171) into the tile-packed fp32 state and its 16-bit operand copy.
__global__ void wide_embed_kernel(const float* __restrict__ emb, int vocab, const int32_t* __restrict__ atom_id, int n_atoms,
                                  float* __restrict__ h32, uint8_t* __restrict__ h16) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (tile, piece, row in tile)
  const int r = (int)(idx & 127), p = (int)((idx >> 7) & 31);
  const int row = (int)(idx >> 12) * 128 + r;
  if (row >= n_atoms) return;
  const int id = min(max(__ldg(atom_id + row), 0), vocab - 1);
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(emb + (int64_t)id * D + 8 * p));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(emb + (int64_t)id * D + 8 * p + 4));
  uint8_t* h32b = reinterpret_cast<uint8_t*>(h32);
  *reinterpret_cast<float4*>(h32b + tp32_off(row, 2 * p)) = x0;
  *reinterpret_cast<float4*>(h32b + tp32_off(row, 2 * p + 1)) = x1;
  *reinterpret_cast<uint4*>(h16 + tp16_off(row, p)) =
      make_uint4(tc::pack_f16x2(x0.x, x0.y), tc::pack_f16x2(x0.z, x0.w), tc::pack_f16x2(x1.x, x1.y), tc::pack_f16x2(x1.z, x1.w));
}

// GlobalSumPool.call (models/layers.py:161-164) from the tile-packed state: one thread per (molecule, 4 columns).
__global__ void wide_pool_kernel(const int32_t* __restrict__ mol_ptr, const int32_t* __restrict__ atom_id, int n_mols,
                                 const float* __restrict__ h32, float* __restrict__ pooled) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = (int)(idx & 63);
  const int m = (int)(idx >> 6);
  if (m >= n_mols) return;
  const int a0 = __ldg(mol_ptr + m), a1 = __ldg(mol_ptr + m + 1);
  const uint8_t* h32b = reinterpret_cast<const uint8_t*>(h32);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = a0; r < a1; ++r) {
    if (__ldg(atom_id + r) <= 0) continue;
    const float4 x = *reinterpret_cast<const float4*>(h32b + tp32_off(r, q));
    s.x += x.x, s.y += x.y, s.z += x.z, s.w += x.w;
  }
  *reinterpret_cast<float4*>(pooled + (int64_t)m * D + 4 * q) = s;
}

// Weight pre-pack (once per weight update): every K = 64 slice of an operand is one contiguous 32 KB image of the
// canonical K-major layout [piece c of 8][n of 256][8 halfs].
//   Wc[m*8+k][n] = bond_transform[k][n][m]  (models/layers.py:94-98,108-112: A[l,m], l = output)
//   gates: B[kk][n] = kernel[kk][n], kernel (2d, d) (rows [0,d) multiply h or r*h, rows [d,2d) multiply agg)
__global__ void wide_pack_kernel(const float* __restrict__ bt, imp_gru_weights_t w, uint8_t* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  __half* o = reinterpret_cast<__half*>(out);
  const int64_t n_wc = (int64_t)D * KB * D, n_g = (int64_t)2 * D * D;
  if (idx < n_wc) {
    const int i = (int)(idx & 7), n = (int)((idx >> 3) & 255), c = (int)((idx >> 11) & 7), j = (int)(idx >> 14);
    const int kk = j * KC + c * 8 + i, m = kk >> 3, k = kk & 7;
    o[idx] = __float2half_rn(bt[((int64_t)k * D + n) * D + m]);
  } else if (idx < n_wc + 3 * n_g) {
    const int64_t t = idx - n_wc;
    const int g = (int)(t / n_g);
    const int64_t u = t % n_g;
    const int i = (int)(u & 7), n = (int)((u >> 3) & 255), c = (int)((u >> 11) & 7), j = (int)(u >> 14);
    const int kk = j * KC + c * 8 + i;
    const float* W = g == 0 ? w.Wz : g == 1 ? w.Wr : w.Wh;
    o[idx] = __float2half_rn(W[(int64_t)kk * D + n]);
  } else if (idx < n_wc + 3 * n_g + 5 * D) {
    const int t = (int)(idx - n_wc - 3 * n_g);
    const float* v = t < D ? w.bz : t < 2 * D ? w.br : t < 3 * D ? w.bh : t < 4 * D ? w.gamma : w.beta;
    reinterpret_cast<float*>(out + OFF_VEC)[t] = v[t % D];
  }
}

}  // namespace wide
}  // namespace imp

using namespace imp;

extern "C" int64_t imp_wide_pack_bytes(int32_t d, int32_t bond_dim) {
  if (d != wide::D || bond_dim != wide::KB) return IMP_ERR_DIM;
  return wide::PACK_BYTES;
}

extern "C" int imp_wide_pack(const float* d_bond_transform, const imp_gru_weights_t* w, int32_t d, int32_t bond_dim,
                             void* d_packed, void* stream) {
  IMP_REQUIRE(d == wide::D && bond_dim == wide::KB, IMP_ERR_DIM, "imp_wide_pack: atom_dim %d / bond_dim %d (supported: 256 / 8)", d,
              bond_dim);
  IMP_REQUIRE(d_bond_transform && w && d_packed, IMP_ERR_ARG, "imp_wide_pack: null pointer");
  const int64_t n = (int64_t)wide::D * wide::KB * wide::D + 3 * (int64_t)2 * wide::D * wide::D + 5 * wide::D;
  wide::wide_pack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_bond_transform, *w,
                                                                                         reinterpret_cast<uint8_t*>(d_packed));
  IMP_LAUNCH_CHECK();
  return 0;
}

// h32 | h16 | agg16 | z16 | rh16, rows padded to a super-tile
extern "C" int64_t imp_wide_workspace_bytes(int32_t n_atoms, int32_t d) {
  if (d != wide::D || n_atoms < 0) return IMP_ERR_DIM;
  const int64_t rows = ceil_div(n_atoms, wide::ST_ROWS) * wide::ST_ROWS;
  return rows * wide::D * (4 + 4 * 2);
}

namespace {

int wide_check(const imp_graph_t* g, int32_t d, int32_t bond_dim, int32_t flags, const void* ws, const char* what) {
  IMP_REQUIRE(g, IMP_ERR_ARG, "%s: graph is null", what);
  IMP_REQUIRE(d == wide::D && bond_dim == wide::KB, IMP_ERR_DIM, "%s: atom_dim %d / bond_dim %d (supported: 256 / 8)", what, d, bond_dim);
  IMP_REQUIRE(flags & IMP_TC_FP16, IMP_ERR_ARG, "%s: IEEE-half operands only (IMP_TC_FP16)", what);
  IMP_REQUIRE(g->bond_vocab <= wide::MAX_BOND_VOCAB, IMP_ERR_ARG, "%s: bond vocabulary > 256", what);
  IMP_REQUIRE(ws || g->n_atoms == 0, IMP_ERR_ARG, "%s: workspace is null", what);
  return 0;
}

void wide_args(const imp_graph_t* g, void* d_workspace, wide::Args* a) {
  const int64_t rows = ceil_div(g->n_atoms, wide::ST_ROWS) * wide::ST_ROWS;
  uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
  a->n_atoms = g->n_atoms, a->n_cat = g->n_cat_atoms;
  a->row_ptr = g->row_ptr, a->col_src = g->col_src, a->edge_bm = g->edge_bm;
  a->bond_emb = nullptr, a->bond_vocab = g->bond_vocab;
  a->packed[0] = a->packed[1] = nullptr;
  a->h32 = reinterpret_cast<float*>(ws);
  a->h16 = ws + rows * wide::D * 4;
  a->agg16 = a->h16 + rows * wide::D * 2;
  a->z16 = a->agg16 + rows * wide::D * 2;
  a->rh16 = a->z16 + rows * wide::D * 2;
  a->eps = 0.f, a->precise = 0;
  a->timeline = nullptr;
}

template <int MODE>
int wide_launch(const wide::Args& a, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    IMP_CUDA(cudaFuncSetAttribute(wide::wide_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::SMEM_BYTES));
    attr_done = true;
  }
  int dev = 0, sms = 148;
  IMP_CUDA(cudaGetDevice(&dev));
  IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int items = wide::n_items(a.n_atoms, a.n_cat) * (MODE == wide::MODE_GRU1 ? 2 : 1);
  if (items == 0) return 0;
  wide::wide_gemm_kernel<MODE><<<items < sms ? items : sms, wide::THREADS, wide::SMEM_BYTES, st>>>(a);
  IMP_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int imp_wide_embed(const imp_graph_t* g, const float* d_atom_emb, int32_t atom_vocab, int32_t d, void* d_workspace,
                              void* stream) {
  if (int rc = wide_check(g, d, wide::KB, IMP_TC_FP16, d_workspace, "imp_wide_embed")) return rc;
  IMP_REQUIRE(atom_vocab > 0 && (d_atom_emb || g->n_atoms == 0), IMP_ERR_ARG, "imp_wide_embed: embedding table missing");
  if (g->n_atoms == 0) return 0;
  wide::Args a;
  wide_args(g, d_workspace, &a);
  const int64_t n = ceil_div(g->n_atoms, 128) * 128 * 32;
  wide::wide_embed_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_atom_emb, atom_vocab, g->atom_id, g->n_atoms,
                                                                                         a.h32, a.h16);
  IMP_LAUNCH_CHECK();
  return 0;
}

static int wide_part(int part, const imp_graph_t* g, const float* d_bond_emb, int32_t d, int32_t bond_dim, const void* d_packed_cat,
                     const void* d_packed_an, float eps, int32_t flags, void* d_workspace, void* stream, const char* what) {
  if (int rc = wide_check(g, d, bond_dim, flags, d_workspace, what)) return rc;
  if (g->n_atoms == 0) return 0;
  IMP_REQUIRE(d_packed_cat && d_packed_an && (part != 0 || d_bond_emb), IMP_ERR_ARG, "%s: null pointer", what);
  wide::Args a;
  wide_args(g, d_workspace, &a);
  a.bond_emb = d_bond_emb;
  a.packed[0] = reinterpret_cast<const uint8_t*>(d_packed_cat), a.packed[1] = reinterpret_cast<const uint8_t*>(d_packed_an);
  a.eps = eps, a.precise = (flags & IMP_TC_PRECISE_EPILOGUE) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (part == 0) return wide_launch<wide::MODE_MSG>(a, st);
  if (part == 1) return wide_launch<wide::MODE_GRU1>(a, st);
  return wide_launch<wide::MODE_GRU2>(a, st);
}

extern "C" int imp_wide_message(const imp_graph_t* g, const float* d_bond_emb, int32_t d, int32_t bond_dim, const void* d_packed_cat,
                                const void* d_packed_an, int32_t flags, void* d_workspace, void* stream) {
  return wide_part(0, g, d_bond_emb, d, bond_dim, d_packed_cat, d_packed_an, 0.f, flags, d_workspace, stream, "imp_wide_message");
}
extern "C" int imp_wide_gates(const imp_graph_t* g, int32_t d, const void* d_packed_cat, const void* d_packed_an, int32_t flags,
                              void* d_workspace, void* stream) {
  return wide_part(1, g, nullptr, d, wide::KB, d_packed_cat, d_packed_an, 0.f, flags, d_workspace, stream, "imp_wide_gates");
}
extern "C" int imp_wide_candidate(const imp_graph_t* g, int32_t d, const void* d_packed_cat, const void* d_packed_an, float eps,
                                  int32_t flags, void* d_workspace, void* stream) {
  return wide_part(2, g, nullptr, d, wide::KB, d_packed_cat, d_packed_an, eps, flags, d_workspace, stream, "imp_wide_candidate");
}

static long long* g_wide_timeline = nullptr;
// Debug aid, NOT part of the public ABI (include/imp_b200.h does not declare it; tools/wide_timeline.py binds it by name):
// when set to a device buffer of 16 x 8 int64, CTA 0 of imp_wide_gated_update records clock64 at its phase boundaries for
// its first 16 tiles; NULL (default) disables it.
extern "C" void imp_debug_wide_timeline(void* d_buf) { g_wide_timeline = reinterpret_cast<long long*>(d_buf); }

extern "C" int imp_wide_gated_update(const imp_graph_t* g, int32_t d, const void* d_packed_cat, const void* d_packed_an, float eps,
                                     int32_t flags, void* d_workspace, void* stream) {
  if (int rc = wide_check(g, d, wide::KB, flags, d_workspace, "imp_wide_gated_update")) return rc;
  if (g->n_atoms == 0) return 0;
  IMP_REQUIRE(d_packed_cat && d_packed_an, IMP_ERR_ARG, "imp_wide_gated_update: null pointer");
  wide::Args a;
  wide_args(g, d_workspace, &a);
  a.packed[0] = reinterpret_cast<const uint8_t*>(d_packed_cat), a.packed[1] = reinterpret_cast<const uint8_t*>(d_packed_an);
  a.eps = eps, a.precise = (flags & IMP_TC_PRECISE_EPILOGUE) ? 1 : 0;
  a.timeline = g_wide_timeline;
  static bool attr_done = false;
  if (!attr_done) {
    IMP_CUDA(cudaFuncSetAttribute(wide::wide_gru_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::G_SMEM_BYTES));
    IMP_CUDA(cudaFuncSetAttribute(wide::wide_gru_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::G_SMEM_BYTES));
    IMP_CUDA(cudaFuncSetAttribute(wide::wide_gru_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::G_SMEM_BYTES));
    IMP_CUDA(cudaFuncSetAttribute(wide::wide_gru_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide::G_SMEM_BYTES));
    attr_done = true;
  }
  int dev = 0, sms = 148;
  IMP_CUDA(cudaGetDevice(&dev));
  IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & IMP_TC_WIDE_NO_CLUSTER)) {
    // clusters of two CTAs share the weight stream (multicast): the grid is the number of clusters that can be co-resident
    const int items = wide::n_gitems<2>(a.n_atoms, a.n_cat);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(2, 1, 1), cfg.blockDim = dim3(wide::G_THREADS, 1, 1), cfg.dynamicSmemBytes = wide::G_SMEM_BYTES, cfg.stream = st;
    cfg.attrs = attr, cfg.numAttrs = 1;
    static int max_clusters = -1;
    if (max_clusters < 0) {
      int n = 0;
      cfg.gridDim = dim3(sms & ~1, 1, 1);
      IMP_CUDA(cudaOccupancyMaxActiveClusters(&n, wide::wide_gru_kernel<false, 2>, &cfg));
      max_clusters = n > 0 ? n : 1;
    }
    const int clusters = items / 2 < max_clusters ? items / 2 : max_clusters;
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    if (a.precise)
      IMP_CUDA(cudaLaunchKernelEx(&cfg, wide::wide_gru_kernel<true, 2>, a));
    else
      IMP_CUDA(cudaLaunchKernelEx(&cfg, wide::wide_gru_kernel<false, 2>, a));
    return 0;
  }
  const int items = wide::n_gitems<1>(a.n_atoms, a.n_cat);
  if (a.precise)
    wide::wide_gru_kernel<true, 1><<<items < sms ? items : sms, wide::G_THREADS, wide::G_SMEM_BYTES, st>>>(a);
  else
    wide::wide_gru_kernel<false, 1><<<items < sms ? items : sms, wide::G_THREADS, wide::G_SMEM_BYTES, st>>>(a);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_wide_pool(const imp_graph_t* g, int32_t d, const void* d_workspace, float* d_pooled, void* stream) {
  if (int rc = wide_check(g, d, wide::KB, IMP_TC_FP16, d_workspace, "imp_wide_pool")) return rc;
  if (g->n_pairs == 0) return 0;
  IMP_REQUIRE(d_pooled, IMP_ERR_ARG, "imp_wide_pool: output is null");
  wide::wide_pool_kernel<<<(unsigned)ceil_div((int64_t)2 * g->n_pairs * 64, 256), 256, 0, (cudaStream_t)stream>>>(
      g->mol_ptr, g->atom_id, 2 * g->n_pairs, reinterpret_cast<const float*>(d_workspace), d_pooled);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_mpnn_forward_wide(const imp_graph_t* g, const float* d_atom_emb, int32_t atom_vocab, const float* d_bond_emb,
                                     int32_t d, int32_t bond_dim, int32_t num_steps, const void* d_packed, float eps,
                                     int32_t flags, void* d_workspace, float* d_pooled, void* stream) {
  if (int rc = wide_check(g, d, bond_dim, flags, d_workspace, "imp_mpnn_forward_wide")) return rc;
  IMP_REQUIRE(num_steps >= 0 && (d_packed || num_steps == 0), IMP_ERR_ARG, "imp_mpnn_forward_wide: num_steps / packed weights");
  if (g->n_pairs == 0) return 0;
  if (int rc = imp_wide_embed(g, d_atom_emb, atom_vocab, d, d_workspace, stream)) return rc;
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(d_packed);
  for (int s = 0; s < num_steps; ++s) {
    const void *pc = pk + (int64_t)s * wide::PACK_BYTES, *pa = pk + (int64_t)(num_steps + s) * wide::PACK_BYTES;
    if (int rc = imp_wide_message(g, d_bond_emb, d, bond_dim, pc, pa, flags, d_workspace, stream)) return rc;
    if (flags & IMP_TC_WIDE_SPLIT_GRU) {
      if (int rc = imp_wide_gates(g, d, pc, pa, flags, d_workspace, stream)) return rc;
      if (int rc = imp_wide_candidate(g, d, pc, pa, eps, flags, d_workspace, stream)) return rc;
    } else if (int rc = imp_wide_gated_update(g, d, pc, pa, eps, flags, d_workspace, stream)) {
      return rc;
    }
  }
  return imp_wide_pool(g, d, d_workspace, d_pooled, stream);
}
