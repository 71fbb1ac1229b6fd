// fp32 SIMT forward kernels of the MPNN hot path (the 1e-5 parity path).
//
//   K1 embed_atoms        Embedding(atom)                       train_viscosity.py:163,171
//   K2 bond_table         Embedding(bond) + tf.tensordot        models/layers.py:108
//   K3+K4 message_agg     BondMatrixMessage o Reduce (fused)    models/layers.py:100-117, 57-83
//   K3 edge_messages      BondMatrixMessage                     models/layers.py:100-117
//   K4 segment_sum        Reduce                                models/layers.py:57-83
//   K5 gated_update       GatedUpdate                           models/layers.py:142-156
//   K6 pool_head          GlobalSumPool + Dense/mix/head        models/layers.py:161-164, 10-42;
//                                                               train_viscosity.py:189-214
//
// All arithmetic is plain fp32 FMA with full-precision expf / tanhf / sqrtf so that the result stays
// within 1e-5 of the fp64 oracle.  Summation orders are fixed (CSR order), so every kernel is
// deterministic run to run -- unlike the reference's scatter_nd on a GPU.
#include <math.h>

#include "common.cuh"

namespace imp {

// ----------------------------------------------------------------------------------------- K1
__global__ void embed_atoms_kernel(const float4* __restrict__ emb, const int* __restrict__ atom_id,
                                   int64_t total4, int d4, int atom_vocab, float4* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int64_t v = i / d4;
  int c = (int)(i - v * d4);
  int id = atom_id[v];
  id = min(max(id, 0), atom_vocab - 1);  // ids are validated by the packer; never read out of bounds
  out[i] = emb[(int64_t)id * d4 + c];
}

// ----------------------------------------------------------------------------------------- K2
// table[t][v][j] = sum_k bond_emb[v][k] * W_t[k][j],  j = l*d + m in [0, d*d).
// One CTA: all V_b rows x 64 consecutive columns of one table; bond_emb chunk + W chunk staged in smem.
struct TablePtrs {
  const float* W[IMP_MAX_TABLES];
  float* table[IMP_MAX_TABLES];
  float* table_il[IMP_MAX_TABLES];
  float* table_ilT[IMP_MAX_TABLES];  // lane-interleaved layout of the TRANSPOSED matrices (backward: dh = T^T g)
};

constexpr int K2_COLS = 64;
constexpr int K2_KC = 32;
constexpr int K2_MAXV = 128;  // bond vocabulary rows handled per CTA pass

__global__ void __launch_bounds__(256) bond_table_kernel(const float* __restrict__ bond_emb, int V, int K, int d,
                                                          TablePtrs ptrs) {
  __shared__ float sA[K2_MAXV][K2_KC + 1];
  __shared__ float sB[K2_KC][K2_COLS];
  const int t = blockIdx.y;
  const int col0 = blockIdx.x * K2_COLS;
  const int dd = d * d;
  const float* __restrict__ W = ptrs.W[t];
  const int tx = threadIdx.x % K2_COLS;  // column within the tile
  const int ty = threadIdx.x / K2_COLS;  // 0..3, row phase
  for (int v0 = 0; v0 < V; v0 += K2_MAXV) {
    const int nv = min(K2_MAXV, V - v0);
    float acc[K2_MAXV / 4];
#pragma unroll
    for (int i = 0; i < K2_MAXV / 4; ++i) acc[i] = 0.f;
    for (int k0 = 0; k0 < K; k0 += K2_KC) {
      const int nk = min(K2_KC, K - k0);
      for (int i = threadIdx.x; i < nv * K2_KC; i += blockDim.x) {
        int r = i / K2_KC, c = i % K2_KC;
        sA[r][c] = c < nk ? bond_emb[(int64_t)(v0 + r) * K + k0 + c] : 0.f;
      }
      for (int i = threadIdx.x; i < K2_KC * K2_COLS; i += blockDim.x) {
        int r = i / K2_COLS, c = i % K2_COLS;
        sB[r][c] = (r < nk && col0 + c < dd) ? W[(int64_t)(k0 + r) * dd + col0 + c] : 0.f;
      }
      __syncthreads();
      for (int k = 0; k < K2_KC; ++k) {
        float b = sB[k][tx];
#pragma unroll
        for (int i = 0; i < K2_MAXV / 4; ++i) acc[i] = fmaf(sA[ty + 4 * i][k], b, acc[i]);
      }
      __syncthreads();
    }
    const int j = col0 + tx;
    if (j < dd) {
      const int l = j / d, m = j % d;
#pragma unroll
      for (int i = 0; i < K2_MAXV / 4; ++i) {
        int v = v0 + ty + 4 * i;
        if (ty + 4 * i < nv) {
          if (ptrs.table[t]) ptrs.table[t][(int64_t)v * dd + j] = acc[i];
          if (ptrs.table_il[t])  // [v][m/4][l][m%4]
            ptrs.table_il[t][(((int64_t)v * (d / 4) + m / 4) * d + l) * 4 + (m & 3)] = acc[i];
          if (ptrs.table_ilT[t])  // [v][l/4][m][l%4]
            ptrs.table_ilT[t][(((int64_t)v * (d / 4) + l / 4) * d + m) * 4 + (l & 3)] = acc[i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ K3 + K4
// One thread per (destination atom v, output component l).  The table is read in the lane-interleaved
// layout so that the D threads of a destination read consecutive float4s; h[src] is a warp-uniform
// (broadcast) 16-byte read.
template <int D>
__device__ __forceinline__ float bond_matvec_row(const float4* __restrict__ trow, const float4* __restrict__ hrow) {
  float s = 0.f;
#pragma unroll
  for (int m4 = 0; m4 < D / 4; ++m4) {
    float4 tv = __ldg(trow + m4 * D);
    float4 hv = __ldg(hrow + m4);
    s = fmaf(tv.x, hv.x, s);
    s = fmaf(tv.y, hv.y, s);
    s = fmaf(tv.z, hv.z, s);
    s = fmaf(tv.w, hv.w, s);
  }
  return s;
}

template <int D>
__global__ void __launch_bounds__(256) message_agg_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_src,
                                                           const int* __restrict__ edge_bm, const float* __restrict__ h,
                                                           const float* __restrict__ tab_cat,
                                                           const float* __restrict__ tab_an, int n_atoms, int n_cat,
                                                           float* __restrict__ agg, int accumulate) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = (int)(t / D), l = (int)(t % D);
  if (v >= n_atoms) return;
  const float* tab = v < n_cat ? tab_cat : tab_an;
  float acc = 0.f;
  const int e1 = row_ptr[v + 1];
  for (int e = row_ptr[v]; e < e1; ++e) {
    const int src = col_src[e];
    const int bm = edge_bm[e];
    const float4* trow = reinterpret_cast<const float4*>(tab + (int64_t)(bm & 0xFFFF) * D * D) + l;
    const float4* hrow = reinterpret_cast<const float4*>(h + (int64_t)src * D);
    acc = fmaf((float)(bm >> 16), bond_matvec_row<D>(trow, hrow), acc);
  }
  agg[(int64_t)v * D + l] = accumulate ? agg[(int64_t)v * D + l] + acc : acc;
}

// K3 alone: messages bucket by bucket, written at the entry's CSR position.  Uses the row-major table.
template <int D>
__global__ void __launch_bounds__(256) edge_messages_kernel(const int* __restrict__ bucket_perm, const int* __restrict__ col_src,
                                                             const int* __restrict__ edge_bm, const float* __restrict__ h,
                                                             const float* __restrict__ tab_cat,
                                                             const float* __restrict__ tab_an, int n_unique,
                                                             const int* __restrict__ first_anion_slot_ptr,
                                                             float* __restrict__ msg) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int i = (int)(t / D), l = (int)(t % D);
  if (i >= n_unique) return;
  const int first_anion_slot = __ldg(first_anion_slot_ptr);  // = bucket_ptr[V_b]: cation buckets come first
  const int e = bucket_perm[i];
  const int bm = edge_bm[e];
  const float* tab = (i < first_anion_slot ? tab_cat : tab_an) + (int64_t)(bm & 0xFFFF) * D * D + (int64_t)l * D;
  const float4* trow = reinterpret_cast<const float4*>(tab);
  const float4* hrow = reinterpret_cast<const float4*>(h + (int64_t)col_src[e] * D);
  float s = 0.f;
#pragma unroll
  for (int m4 = 0; m4 < D / 4; ++m4) {
    float4 tv = __ldg(trow + m4);
    float4 hv = __ldg(hrow + m4);
    s = fmaf(tv.x, hv.x, s);
    s = fmaf(tv.y, hv.y, s);
    s = fmaf(tv.z, hv.z, s);
    s = fmaf(tv.w, hv.w, s);
  }
  msg[(int64_t)e * D + l] = (float)(bm >> 16) * s;
}

// K3 for atom_dim 32, bucket-grouped (the fp32 twin of csrc/msg_tc.cu): a CTA takes <= 128 consecutive slots of ONE
// (tower, bond) bucket, so the 4 KB bond matrix T[b] is staged in shared memory once and read as warp-broadcast rows by every
// entry of the chunk (the CSR-order kernels re-read a different matrix per entry through L1, which is what bounds them).
// Source rows are gathered 8 lanes per 128-byte row into a padded tile, each thread then owns one entry:
//   forward      msg[e][l] = mult * sum_m T[l][m] * x[src_e][m]                (BondMatrixMessage.call, models/layers.py:108-112)
//   TRANSPOSED   msg[e][m] = mult * sum_l T[l][m] * x[src_e][l]                (its backward with respect to the atom states)
// and rows go back through the tile so that stores are whole lines.  Exact fp32, deterministic.
constexpr int GM_CHUNK = 128;
constexpr int GM_LD = 36;  // padded row: 16-byte aligned, conflict-free float4 accesses of consecutive rows

__global__ void gm_chunk_scan_kernel(const int* __restrict__ bucket_ptr, int n_buckets, int* __restrict__ chunk_ptr) {
  if (threadIdx.x == 0) {
    int c = 0;
    for (int b = 0; b < n_buckets; ++b) {
      chunk_ptr[b] = c;
      c += (bucket_ptr[b + 1] - bucket_ptr[b] + GM_CHUNK - 1) / GM_CHUNK;
    }
    chunk_ptr[n_buckets] = c;
  }
}

template <bool TRANSPOSED>
__global__ void __launch_bounds__(GM_CHUNK) grouped_msg_f32_kernel(const int* __restrict__ bucket_ptr, const int* __restrict__ chunk_ptr,
                                                                   int n_buckets, int bond_vocab, const int* __restrict__ bucket_perm,
                                                                   const int* __restrict__ col_src, const int* __restrict__ edge_bm,
                                                                   const float* __restrict__ x, const float* __restrict__ tab_cat,
                                                                   const float* __restrict__ tab_an, float* __restrict__ msg) {
  constexpr int D = 32;
  __shared__ __align__(16) float sT[D * D];
  __shared__ __align__(16) float stg[GM_CHUNK * GM_LD];
  const int chunk = blockIdx.x;
  if (chunk >= __ldg(chunk_ptr + n_buckets)) return;
  int lo = 0, hi = n_buckets - 1;
  while (lo < hi) {  // bucket of this chunk: last b with chunk_ptr[b] <= chunk
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
  }
  const int b = lo;
  const int slot0 = __ldg(bucket_ptr + b) + (chunk - __ldg(chunk_ptr + b)) * GM_CHUNK;
  const int n = min(GM_CHUNK, __ldg(bucket_ptr + b + 1) - slot0);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float4* tb = reinterpret_cast<const float4*>((b < bond_vocab ? tab_cat + (int64_t)b * D * D : tab_an + (int64_t)(b - bond_vocab) * D * D));
  reinterpret_cast<float4*>(sT)[t] = __ldg(tb + t);
  reinterpret_cast<float4*>(sT)[t + GM_CHUNK] = __ldg(tb + t + GM_CHUNK);
  int e = -1, src = 0;
  float mult = 0.f;
  if (t < n) {
    e = __ldg(bucket_perm + slot0 + t);
    mult = (float)((unsigned)__ldg(edge_bm + e) >> 16);
    src = __ldg(col_src + e);
  }
  const int g = lane >> 3, q = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + g;
    const int rs = __shfl_sync(0xffffffffu, src, r), re = __shfl_sync(0xffffffffu, e, r);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (re >= 0) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)rs * D) + q);
    *reinterpret_cast<float4*>(stg + (warp * 32 + r) * GM_LD + 4 * q) = v;
  }
  __syncthreads();  // sT complete, and this warp's rows of stg
  float xin[D], out[D];
#pragma unroll
  for (int c = 0; c < D / 4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(stg + t * GM_LD + 4 * c);
    xin[4 * c] = v.x, xin[4 * c + 1] = v.y, xin[4 * c + 2] = v.z, xin[4 * c + 3] = v.w;
  }
  if (!TRANSPOSED) {
#pragma unroll
    for (int l = 0; l < D; ++l) {
      const float4* row = reinterpret_cast<const float4*>(sT + l * D);
      float a = 0.f;
#pragma unroll
      for (int m4 = 0; m4 < D / 4; ++m4) {
        const float4 w = row[m4];
        a = fmaf(w.x, xin[4 * m4], a), a = fmaf(w.y, xin[4 * m4 + 1], a), a = fmaf(w.z, xin[4 * m4 + 2], a), a = fmaf(w.w, xin[4 * m4 + 3], a);
      }
      out[l] = a;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) out[c] = 0.f;
#pragma unroll
    for (int l = 0; l < D; ++l) {
      const float4* row = reinterpret_cast<const float4*>(sT + l * D);
#pragma unroll
      for (int m4 = 0; m4 < D / 4; ++m4) {
        const float4 w = row[m4];
        out[4 * m4] = fmaf(w.x, xin[l], out[4 * m4]), out[4 * m4 + 1] = fmaf(w.y, xin[l], out[4 * m4 + 1]);
        out[4 * m4 + 2] = fmaf(w.z, xin[l], out[4 * m4 + 2]), out[4 * m4 + 3] = fmaf(w.w, xin[l], out[4 * m4 + 3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < D / 4; ++c)
    *reinterpret_cast<float4*>(stg + t * GM_LD + 4 * c) = make_float4(mult * out[4 * c], mult * out[4 * c + 1], mult * out[4 * c + 2], mult * out[4 * c + 3]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + g;
    const int re = __shfl_sync(0xffffffffu, e, r);
    const float4 sr = *reinterpret_cast<const float4*>(stg + (warp * 32 + r) * GM_LD + 4 * q);
    if (re >= 0) reinterpret_cast<float4*>(msg + (int64_t)re * D)[q] = sr;
  }
}

// K4: contiguous segment sum over CSR rows, one thread per (v, float4 column).
__global__ void segment_sum_kernel(const int* __restrict__ row_ptr, const float4* __restrict__ msg, int n_atoms, int d4,
                                   float4* __restrict__ agg, int accumulate) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = (int)(t / d4), c = (int)(t % d4);
  if (v >= n_atoms) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const int e1 = row_ptr[v + 1];
  for (int e = row_ptr[v]; e < e1; e += 4) {  // four rows in flight; summed in entry order (+0 for the missing ones)
    float4 m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) m[u] = e + u < e1 ? __ldg(msg + (int64_t)(e + u) * d4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a.x += m[u].x;
      a.y += m[u].y;
      a.z += m[u].z;
      a.w += m[u].w;
    }
  }
  if (accumulate) {
    const float4 o = agg[(int64_t)v * d4 + c];
    a.x += o.x, a.y += o.y, a.z += o.z, a.w += o.w;
  }
  agg[(int64_t)v * d4 + c] = a;
}

// ----------------------------------------------------------------------------------------- K5
// One thread per atom, 128 atoms per CTA, one tower per CTA.  The three (2D x D) kernels live in shared
// memory and are read as warp-uniform float4 broadcasts; the thread's own h / agg / r*h rows sit in
// padded (conflict-free) shared tiles.  All gate math, LayerNorm and the residual are thread-local.
constexpr int K5_TILE = 128;

template <int D>
struct K5Smem {
  float Wz[2 * D * D], Wr[2 * D * D], Wh[2 * D * D];
  float bz[D], br[D], bh[D], gamma[D], beta[D];
  float hs[K5_TILE][D + 1], as[K5_TILE][D + 1], xs[K5_TILE][D + 1];
};

template <int D>
__device__ __forceinline__ void k5_dense(float (&acc)[D], const float* __restrict__ W /* smem [2D][D] */,
                                         const float* __restrict__ bias, const float* __restrict__ x0 /* own row, D */,
                                         const float* __restrict__ x1 /* own row, D */) {
#pragma unroll
  for (int j = 0; j < D; ++j) acc[j] = bias[j];
#pragma unroll 4
  for (int k = 0; k < D; ++k) {
    const float x = x0[k];
    const float4* w = reinterpret_cast<const float4*>(W + k * D);
#pragma unroll
    for (int j4 = 0; j4 < D / 4; ++j4) {
      float4 wv = w[j4];
      acc[4 * j4 + 0] = fmaf(x, wv.x, acc[4 * j4 + 0]);
      acc[4 * j4 + 1] = fmaf(x, wv.y, acc[4 * j4 + 1]);
      acc[4 * j4 + 2] = fmaf(x, wv.z, acc[4 * j4 + 2]);
      acc[4 * j4 + 3] = fmaf(x, wv.w, acc[4 * j4 + 3]);
    }
  }
#pragma unroll 4
  for (int k = 0; k < D; ++k) {
    const float x = x1[k];
    const float4* w = reinterpret_cast<const float4*>(W + (D + k) * D);
#pragma unroll
    for (int j4 = 0; j4 < D / 4; ++j4) {
      float4 wv = w[j4];
      acc[4 * j4 + 0] = fmaf(x, wv.x, acc[4 * j4 + 0]);
      acc[4 * j4 + 1] = fmaf(x, wv.y, acc[4 * j4 + 1]);
      acc[4 * j4 + 2] = fmaf(x, wv.z, acc[4 * j4 + 2]);
      acc[4 * j4 + 3] = fmaf(x, wv.w, acc[4 * j4 + 3]);
    }
  }
}

__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int D>
__global__ void __launch_bounds__(K5_TILE) gated_update_kernel(const float* __restrict__ h, const float* __restrict__ agg,
                                                               int n_atoms, int n_cat, int tiles_cat,
                                                               imp_gru_weights_t wc, imp_gru_weights_t wa, float eps,
                                                               float* __restrict__ h_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K5Smem<D>& s = *reinterpret_cast<K5Smem<D>*>(smem_raw);
  const bool is_cat = (int)blockIdx.x < tiles_cat;
  const imp_gru_weights_t& w = is_cat ? wc : wa;
  const int a0 = is_cat ? blockIdx.x * K5_TILE : n_cat + (blockIdx.x - tiles_cat) * K5_TILE;
  const int a_end = is_cat ? n_cat : n_atoms;
  const int rows = min(K5_TILE, a_end - a0);
  const int tid = threadIdx.x;

  for (int i = tid; i < 2 * D * D / 4; i += K5_TILE) {
    reinterpret_cast<float4*>(s.Wz)[i] = __ldg(reinterpret_cast<const float4*>(w.Wz) + i);
    reinterpret_cast<float4*>(s.Wr)[i] = __ldg(reinterpret_cast<const float4*>(w.Wr) + i);
    reinterpret_cast<float4*>(s.Wh)[i] = __ldg(reinterpret_cast<const float4*>(w.Wh) + i);
  }
  for (int i = tid; i < D; i += K5_TILE) {
    s.bz[i] = w.bz[i];
    s.br[i] = w.br[i];
    s.bh[i] = w.bh[i];
    s.gamma[i] = w.gamma[i];
    s.beta[i] = w.beta[i];
  }
  // coalesced tile loads: consecutive threads read consecutive float4s of the [rows, D] slab
  const float4* hg = reinterpret_cast<const float4*>(h + (int64_t)a0 * D);
  const float4* ag = reinterpret_cast<const float4*>(agg + (int64_t)a0 * D);
  for (int i = tid; i < rows * (D / 4); i += K5_TILE) {
    const int r = i / (D / 4), c = (i % (D / 4)) * 4;
    float4 hv = __ldg(hg + i), av = __ldg(ag + i);
    s.hs[r][c] = hv.x, s.hs[r][c + 1] = hv.y, s.hs[r][c + 2] = hv.z, s.hs[r][c + 3] = hv.w;
    s.as[r][c] = av.x, s.as[r][c + 1] = av.y, s.as[r][c + 2] = av.z, s.as[r][c + 3] = av.w;
  }
  __syncthreads();

  if (tid < rows) {
    float acc[D], g[D];
    // r gate -> r * h into the xs tile (own row only; no cross-thread hazard)
    k5_dense<D>(acc, s.Wr, s.br, s.hs[tid], s.as[tid]);
#pragma unroll
    for (int j = 0; j < D; ++j) s.xs[tid][j] = sigmoidf_precise(acc[j]) * s.hs[tid][j];
    // candidate
    k5_dense<D>(g, s.Wh, s.bh, s.xs[tid], s.as[tid]);
#pragma unroll
    for (int j = 0; j < D; ++j) g[j] = tanhf(g[j]);
    // z gate and blend
    k5_dense<D>(acc, s.Wz, s.bz, s.hs[tid], s.as[tid]);
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const float z = sigmoidf_precise(acc[j]);
      const float hj = s.hs[tid][j];
      g[j] = (1.0f - z) * hj + z * g[j];
      mean += g[j];
    }
    mean *= (1.0f / D);
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const float c = g[j] - mean;
      var = fmaf(c, c, var);
    }
    const float inv = 1.0f / sqrtf(var * (1.0f / D) + eps);
#pragma unroll
    for (int j = 0; j < D; ++j) s.xs[tid][j] = (g[j] - mean) * inv * s.gamma[j] + s.beta[j] + s.hs[tid][j];
  }
  __syncthreads();
  float4* og = reinterpret_cast<float4*>(h_out + (int64_t)a0 * D);
  for (int i = tid; i < rows * (D / 4); i += K5_TILE) {
    const int r = i / (D / 4), c = (i % (D / 4)) * 4;
    og[i] = make_float4(s.xs[r][c], s.xs[r][c + 1], s.xs[r][c + 2], s.xs[r][c + 3]);
  }
}

// ----------------------------------------------------------------------------------------- K6
// One warp per ion pair.  Pool: lanes stride over feature columns (coalesced row reads).  The tiny dense
// layers run out of per-warp shared scratch.  Readout weights are staged in shared memory per CTA.
constexpr int K6_WARPS = 8;
constexpr int K6_MAXV = 256;  // max of d, fp, mix, fp2

struct K6Args {
  const int* mol_ptr;
  const int* atom_id;
  const float* h;
  const float* pooled;           // non-null: [2P, d] molecule sums already computed (fused forward); h / mol_ptr unused
  int n_pairs, d, fp, mix, fp2;  // fp2 = 0 -> viscosity head
  int scratch_stride;
  imp_readout_weights_t wc, wa;
  const float *W1, *b1, *W2, *b2;  // viscosity: W1 = [mix,3] head, b1 = [3]; mp: W1 [mix,fp2], W2 [fp2,1]
  const float* T;
  float* out;
  float* aux;
};

__device__ __forceinline__ float softplusf_precise(float x) {
  return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x));
}

// The readout of the 1e-5 path accumulates in DOUBLE: a head output of magnitude ~1 is a difference of ~20 terms of
// magnitude ~25 (pooled sums of 25 atoms), so fp32 accumulation alone already costs ~1e-5 relative -- the whole north-star
// budget (measured: the reference's own fp32 arithmetic sits at 0.97e-5 on configs[0]).  The work is negligible
// (~6 kFLOP per pair).
__device__ void k6_dense(const float* __restrict__ W /* smem [n_in][n_out] */, const float* __restrict__ b,
                         const double* __restrict__ x /* smem [n_in] */, double* __restrict__ y, int n_in, int n_out,
                         bool relu, int lane) {
  for (int o = lane; o < n_out; o += 32) {
    double a = (double)b[o];
    for (int k = 0; k < n_in; ++k) a = fma(x[k], (double)W[k * n_out + o], a);
    y[o] = relu ? fmax(a, 0.0) : a;
  }
  __syncwarp();
}

__device__ __forceinline__ double softplus_f64(double x) { return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x)); }

// GlobalSumPool alone: one warp per molecule.
__global__ void global_sum_pool_kernel(const int* __restrict__ mol_ptr, const int* __restrict__ atom_id, int n_mols,
                                       const float* __restrict__ h, int d, float* __restrict__ out) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= n_mols) return;
  const int v0 = mol_ptr[m], v1 = mol_ptr[m + 1];
  for (int j = lane; j < d; j += 32) {
    // four rows in flight (the loop is a chain of dependent 128-byte row loads otherwise); rows are added in atom order
    float sacc = 0.f;
    int v = v0;
    for (; v + 4 <= v1; v += 4) {
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = atom_id[v + u] > 0 ? h[(int64_t)(v + u) * d + j] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) sacc += x[u];
    }
    for (; v < v1; ++v)
      if (atom_id[v] > 0) sacc += h[(int64_t)v * d + j];
    out[(int64_t)m * d + j] = sacc;
  }
}

__global__ void __launch_bounds__(K6_WARPS * 32) pool_head_kernel(K6Args a) {
  extern __shared__ __align__(16) float sm[];
  const int d = a.d, fp = a.fp, mix = a.mix, fp2 = a.fp2;
  const int n_head_out = fp2 > 0 ? fp2 : 3;
  // weight staging
  float* Wfp[2] = {sm, sm + d * fp};
  float* p = sm + 2 * d * fp;
  float* bfp[2] = {p, p + fp};
  p += 2 * fp;
  float* Wmx[2] = {p, p + fp * mix};
  p += 2 * fp * mix;
  float* bmx[2] = {p, p + mix};
  p += 2 * mix;
  float* W1 = p;
  p += mix * n_head_out;
  float* b1 = p;
  p += n_head_out;
  float* W2 = p;
  p += (fp2 > 0 ? fp2 : 0);
  const int sv = a.scratch_stride;  // max(d, fp, mix, fp2) rounded up to 32 elements
  p += (reinterpret_cast<uintptr_t>(p) & 7) ? 1 : 0;  // 8-byte alignment of the double scratch
  double* scratch = reinterpret_cast<double*>(p) + (threadIdx.x / 32) * (4 * sv);
  for (int t = 0; t < 2; ++t) {
    const imp_readout_weights_t& w = t == 0 ? a.wc : a.wa;
    for (int i = threadIdx.x; i < d * fp; i += blockDim.x) Wfp[t][i] = w.W_fp[i];
    for (int i = threadIdx.x; i < fp; i += blockDim.x) bfp[t][i] = w.b_fp[i];
    for (int i = threadIdx.x; i < fp * mix; i += blockDim.x) Wmx[t][i] = w.W_mix[i];
    for (int i = threadIdx.x; i < mix; i += blockDim.x) bmx[t][i] = w.b_mix[i];
  }
  for (int i = threadIdx.x; i < mix * n_head_out; i += blockDim.x) W1[i] = a.W1[i];
  for (int i = threadIdx.x; i < n_head_out; i += blockDim.x) b1[i] = a.b1[i];
  if (fp2 > 0)
    for (int i = threadIdx.x; i < fp2; i += blockDim.x) W2[i] = a.W2[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * K6_WARPS + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * K6_WARPS;
  double* pool = scratch;             // [d]
  double* v1 = scratch + sv;     // [fp]
  double* v2 = scratch + 2 * sv; // [mix]
  double* mixed = scratch + 3 * sv;
  const int aux_stride = 2 * d + 2 * fp + mix + (fp2 > 0 ? 0 : 3);
  for (int pair = warp_global; pair < a.n_pairs; pair += n_warps) {
    for (int t = 0; t < 2; ++t) {
      const int m = t * a.n_pairs + pair;
      if (a.pooled) {
        for (int j = lane; j < d; j += 32) pool[j] = (double)a.pooled[(int64_t)m * d + j];
      } else {
        const int v0 = a.mol_ptr[m], v1e = a.mol_ptr[m + 1];
        for (int j = lane; j < d; j += 32) {
          double sacc = 0.0;
          for (int v = v0; v < v1e; ++v)
            if (a.atom_id[v] > 0) sacc += (double)a.h[(int64_t)v * d + j];  // models/layers.py:163 mask
          pool[j] = sacc;
        }
      }
      __syncwarp();
      k6_dense(Wfp[t], bfp[t], pool, v1, d, fp, true, lane);
      k6_dense(Wmx[t], bmx[t], v1, v2, fp, mix, true, lane);
      for (int j = lane; j < mix; j += 32) mixed[j] = t == 0 ? v2[j] : mixed[j] + v2[j];
      if (a.aux) {
        float* ax = a.aux + (int64_t)pair * aux_stride;
        for (int j = lane; j < d; j += 32) ax[t * d + j] = (float)pool[j];
        for (int j = lane; j < fp; j += 32) ax[2 * d + t * fp + j] = (float)v1[j];
      }
      __syncwarp();
    }
    if (a.aux)
      for (int j = lane; j < mix; j += 32) a.aux[(int64_t)pair * aux_stride + 2 * d + 2 * fp + j] = (float)mixed[j];
    if (fp2 == 0) {
      k6_dense(W1, b1, mixed, v1, mix, 3, false, lane);
      if (lane == 0) {
        const double A = v1[0];
        const double B = fmin(fmax(softplus_f64(v1[1]), 0.0), 20.0);
        const double C = fmin(fmax(softplus_f64(v1[2]), 0.1), 50.0);
        const double Ts = (double)a.T[pair] / 100.0;
        a.out[pair] = (float)(A + B / (Ts + C + 1e-6));
        if (a.aux) {
          float* ax = a.aux + (int64_t)pair * aux_stride + 2 * d + 2 * fp + mix;
          ax[0] = (float)A, ax[1] = (float)B, ax[2] = (float)C;
        }
      }
    } else {
      k6_dense(W1, b1, mixed, v1, mix, fp2, true, lane);
      double part = 0.0;
      for (int k = lane; k < fp2; k += 32) part = fma(v1[k], (double)W2[k], part);
      // fixed-order reduction: gather the 32 partials in lane order
      double tot = 0.0;
      for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, part, l);
      if (lane == 0) a.out[pair] = (float)(tot + (double)a.b2[0]);
    }
    __syncwarp();
  }
}


// ----------------------------------------------------------------------------------------- K6, fast form
// Readout on pooled sums for d, fp, mix, fp2 <= 32 (both reference models): ONE THREAD PER ION PAIR.  The five small Dense
// layers (train_viscosity.py:189-204 / train_melting_point.py:173-198) are matrix-vector products of at most 32 x 32 per
// pair; a thread keeps the 32 outputs of a layer in registers, reads its input vector from its own (padded, conflict-free)
// row of shared memory and the weight row W[k][:] as eight broadcast 16-byte reads: 41 instructions per 32 FMAs, no
// cross-lane step, 32 independent accumulation chains per thread (the previous warp-per-pair form with column-resident
// weights needed 160 weight registers per lane -- 8 warps per SM -- and was latency-bound at 0.3 instructions per clock).
// Accumulation order per output: bias, then k ascending -- as pool_head_kernel.
constexpr int R32_THREADS = 128;
constexpr int R32_XS = 33;  // floats per thread row
struct R32Smem {
  float W[5][32 * 32];   // fp_cat, mix_cat, fp_an, mix_an, head (rows k, 32 zero-padded columns)
  float b[5][32];
  float w2[32];
  float x[R32_THREADS * R32_XS];
};
__device__ __forceinline__ void r32_layer(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ xrow, int n_in,
                                          float (&acc)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = b[j];
#pragma unroll 2
  for (int k = 0; k < n_in; ++k) {
    const float xk = xrow[k];
    const float4* w4 = reinterpret_cast<const float4*>(W + k * 32);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 w = w4[j4];
      acc[4 * j4] = fmaf(xk, w.x, acc[4 * j4]), acc[4 * j4 + 1] = fmaf(xk, w.y, acc[4 * j4 + 1]);
      acc[4 * j4 + 2] = fmaf(xk, w.z, acc[4 * j4 + 2]), acc[4 * j4 + 3] = fmaf(xk, w.w, acc[4 * j4 + 3]);
    }
  }
}
__global__ void __launch_bounds__(R32_THREADS) readout32_kernel(K6Args a) {
  extern __shared__ __align__(16) unsigned char r32_raw[];
  R32Smem& s = *reinterpret_cast<R32Smem*>(r32_raw);
  const int d = a.d, fp = a.fp, mix = a.mix, fp2 = a.fp2, nh = fp2 > 0 ? fp2 : 3;
  const int tid = threadIdx.x;
  for (int i = tid; i < 5 * 32 * 32; i += R32_THREADS) {
    const int m = i / 1024, k = (i % 1024) / 32, j = i % 32;
    float v = 0.f;
    if (m == 0 && k < d && j < fp) v = a.wc.W_fp[k * fp + j];
    if (m == 1 && k < fp && j < mix) v = a.wc.W_mix[k * mix + j];
    if (m == 2 && k < d && j < fp) v = a.wa.W_fp[k * fp + j];
    if (m == 3 && k < fp && j < mix) v = a.wa.W_mix[k * mix + j];
    if (m == 4 && k < mix && j < nh) v = a.W1[k * nh + j];
    s.W[m][k * 32 + j] = v;
  }
  for (int i = tid; i < 5 * 32; i += R32_THREADS) {
    const int m = i / 32, j = i % 32;
    float v = 0.f;
    if (m == 0 && j < fp) v = a.wc.b_fp[j];
    if (m == 1 && j < mix) v = a.wc.b_mix[j];
    if (m == 2 && j < fp) v = a.wa.b_fp[j];
    if (m == 3 && j < mix) v = a.wa.b_mix[j];
    if (m == 4 && j < nh) v = a.b1[j];
    s.b[m][j] = v;
  }
  if (tid < 32) s.w2[tid] = (fp2 > 0 && tid < fp2) ? a.W2[tid] : 0.f;
  __syncthreads();
  float* xrow = &s.x[tid * R32_XS];
  float acc[32], mixed[32];
  // a warp stays on 32 consecutive pairs and loops while ANY of them exists (its loads are cooperative); lanes past the end
  // compute on zero rows and store nothing
  for (int pair0w = blockIdx.x * R32_THREADS + (tid & ~31); pair0w < a.n_pairs; pair0w += gridDim.x * R32_THREADS) {
    const int pair = pair0w + (tid & 31);
    const bool live = pair < a.n_pairs;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (d == 32) {
        // the warp's 32 pooled rows are consecutive: 8 lanes per 128-byte row (4 lines per load instruction instead of the 32 that
        // a thread reading its own row touches), into the owners' shared-memory rows
        const int lane = tid & 31, wbase = tid & ~31, g = lane >> 3, qc = lane & 7;
        const int pair0 = pair0w;
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = 4 * it + g;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pair0 + rr < a.n_pairs) v = __ldg(reinterpret_cast<const float4*>(a.pooled + (int64_t)(t * a.n_pairs + pair0 + rr) * 32) + qc);
          float* xr = &s.x[(wbase + rr) * R32_XS + 4 * qc];
          xr[0] = v.x, xr[1] = v.y, xr[2] = v.z, xr[3] = v.w;
        }
        __syncwarp();
      } else {
        const float4* src = reinterpret_cast<const float4*>(a.pooled + (int64_t)(t * a.n_pairs + (live ? pair : 0)) * d);
        for (int c = 0; c < d / 4; ++c) {
          const float4 v = __ldg(src + c);
          xrow[4 * c] = v.x, xrow[4 * c + 1] = v.y, xrow[4 * c + 2] = v.z, xrow[4 * c + 3] = v.w;
        }
      }
      r32_layer(s.W[2 * t], s.b[2 * t], xrow, d, acc);        // Dense(fp_size, relu)
#pragma unroll
      for (int j = 0; j < 32; ++j) xrow[j] = fmaxf(acc[j], 0.f);
      r32_layer(s.W[2 * t + 1], s.b[2 * t + 1], xrow, fp, acc);  // Dense(mixing_size, relu)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v2 = fmaxf(acc[j], 0.f);
        mixed[j] = t == 0 ? v2 : mixed[j] + v2;  // AddTwoTensors / Add
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) xrow[j] = mixed[j];
    r32_layer(s.W[4], s.b[4], xrow, mix, acc);
    if (fp2 == 0) {  // A + B / (T / 100 + C + 1e-6), B and C clipped softplus (models/layers.py:10-42)
      const float B = fminf(fmaxf(softplusf_precise(acc[1]), 0.0f), 20.0f);
      const float Cc = fminf(fmaxf(softplusf_precise(acc[2]), 0.1f), 50.0f);
      if (live) a.out[pair] = acc[0] + B / (__ldg(a.T + pair) / 100.0f + Cc + 1e-6f);
    } else {
      float tot = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) tot = __fadd_rn(tot, __fmul_rn(fmaxf(acc[j], 0.f), s.w2[j]));  // fixed order, no contraction
      if (live) a.out[pair] = tot + a.b2[0];
    }
  }
}

}  // namespace imp

// =========================================================================================== ABI
using namespace imp;

static bool dim_ok(int d) { return d == 8 || d == 16 || d == 32 || d == 64; }                      // thread-per-atom GRU
static bool dim_ok_msg(int d) { return dim_ok(d) || d == 128 || d == 256; }                          // message kernels

extern "C" int imp_embed_atoms(const float* d_atom_emb, int32_t atom_vocab, const int32_t* d_atom_id, int32_t n_atoms,
                               int32_t d, float* d_h0, void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && atom_vocab > 0, IMP_ERR_ARG, "imp_embed_atoms: bad sizes");
  IMP_REQUIRE(d > 0 && d % 4 == 0, IMP_ERR_DIM, "imp_embed_atoms: atom_dim %d must be a multiple of 4", d);
  if (n_atoms == 0) return 0;
  IMP_REQUIRE(d_atom_emb && d_atom_id && d_h0, IMP_ERR_ARG, "imp_embed_atoms: null pointer");
  const int64_t total4 = (int64_t)n_atoms * (d / 4);
  embed_atoms_kernel<<<(unsigned)ceil_div(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(d_atom_emb), d_atom_id, total4, d / 4, atom_vocab, reinterpret_cast<float4*>(d_h0));
  IMP_LAUNCH_CHECK();
  return 0;
}

static int bond_table_any(const float* d_bond_emb, int32_t bond_vocab, int32_t bond_dim, int32_t d, int32_t n_tables,
                          const float* const* h_W, float* const* h_table, float* const* h_table_il, float* const* h_table_ilT,
                          void* stream) {
  IMP_REQUIRE(d_bond_emb && h_W && (h_table || h_table_il || h_table_ilT), IMP_ERR_ARG, "imp_bond_table: null pointer");
  IMP_REQUIRE(n_tables > 0 && n_tables <= IMP_MAX_TABLES, IMP_ERR_ARG, "imp_bond_table: n_tables %d not in 1..%d",
              n_tables, IMP_MAX_TABLES);
  IMP_REQUIRE(bond_vocab > 0 && bond_dim > 0, IMP_ERR_ARG, "imp_bond_table: bad sizes");
  IMP_REQUIRE(d > 0 && d % 4 == 0, IMP_ERR_DIM, "imp_bond_table: atom_dim %d must be a multiple of 4", d);
  TablePtrs p;
  for (int i = 0; i < IMP_MAX_TABLES; ++i) {
    p.W[i] = i < n_tables ? h_W[i] : nullptr;
    p.table[i] = (i < n_tables && h_table) ? h_table[i] : nullptr;
    p.table_il[i] = (i < n_tables && h_table_il) ? h_table_il[i] : nullptr;
    p.table_ilT[i] = (i < n_tables && h_table_ilT) ? h_table_ilT[i] : nullptr;
    IMP_REQUIRE(i >= n_tables || p.W[i], IMP_ERR_ARG, "imp_bond_table: W[%d] is null", i);
  }
  dim3 grid((unsigned)ceil_div((int64_t)d * d, K2_COLS), (unsigned)n_tables);
  bond_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_bond_emb, bond_vocab, bond_dim, d, p);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_bond_table(const float* d_bond_emb, int32_t bond_vocab, int32_t bond_dim, int32_t d, int32_t n_tables,
                              const float* const* h_W, float* const* h_table, float* const* h_table_il, void* stream) {
  return bond_table_any(d_bond_emb, bond_vocab, bond_dim, d, n_tables, h_W, h_table, h_table_il, nullptr, stream);
}

extern "C" int imp_bond_table_train(const float* d_bond_emb, int32_t bond_vocab, int32_t bond_dim, int32_t d, int32_t n_tables,
                                    const float* const* h_W, float* const* h_table_il, float* const* h_table_ilT, void* stream) {
  return bond_table_any(d_bond_emb, bond_vocab, bond_dim, d, n_tables, h_W, nullptr, h_table_il, h_table_ilT, stream);
}

static int check_graph(const imp_graph_t* g, const char* who) {
  IMP_REQUIRE(g, IMP_ERR_ARG, "%s: graph is null", who);
  IMP_REQUIRE(g->n_atoms >= 0 && g->n_unique >= 0 && g->n_pairs >= 0 && g->n_cat_atoms >= 0 &&
                  g->n_cat_atoms <= g->n_atoms && g->bond_vocab > 0,
              IMP_ERR_ARG, "%s: inconsistent graph sizes", who);
  return 0;
}

#define IMP_DISPATCH_D(d, CALL)            \
  switch (d) {                             \
    case 8: { constexpr int D = 8; CALL; } break;   \
    case 16: { constexpr int D = 16; CALL; } break; \
    case 32: { constexpr int D = 32; CALL; } break; \
    case 64: { constexpr int D = 64; CALL; } break; \
    default: break;                        \
  }

#define IMP_DISPATCH_D_MSG(d, CALL)        \
  switch (d) {                             \
    case 8: { constexpr int D = 8; CALL; } break;   \
    case 16: { constexpr int D = 16; CALL; } break; \
    case 32: { constexpr int D = 32; CALL; } break; \
    case 64: { constexpr int D = 64; CALL; } break; \
    case 128: { constexpr int D = 128; CALL; } break; \
    case 256: { constexpr int D = 256; CALL; } break; \
    default: break;                        \
  }

static int message_agg_any(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_il_cat,
                           const float* d_table_il_an, float* d_agg, int accumulate, void* stream);

extern "C" int imp_message_agg(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_il_cat,
                               const float* d_table_il_an, float* d_agg, void* stream) {
  return message_agg_any(g, d_h, d, d_table_il_cat, d_table_il_an, d_agg, 0, stream);
}

extern "C" int imp_message_agg_bwd(const imp_graph_t* g, const float* d_dagg, int32_t d, const float* d_table_ilT_cat,
                                   const float* d_table_ilT_an, float* d_dh, void* stream) {
  return message_agg_any(g, d_dagg, d, d_table_ilT_cat, d_table_ilT_an, d_dh, 1, stream);
}

static int message_agg_any(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_il_cat,
                           const float* d_table_il_an, float* d_agg, int accumulate, void* stream) {
  if (int rc = check_graph(g, "imp_message_agg")) return rc;
  IMP_REQUIRE(dim_ok_msg(d), IMP_ERR_DIM, "imp_message_agg: atom_dim %d not in {8,16,32,64,128,256} (fp32 path)", d);
  if (g->n_atoms == 0) return 0;
  IMP_REQUIRE(d_h && d_table_il_cat && d_table_il_an && d_agg && g->row_ptr && (g->n_unique == 0 || (g->col_src && g->edge_bm)),
              IMP_ERR_ARG, "imp_message_agg: null pointer");
  const int64_t threads = (int64_t)g->n_atoms * d;
  const unsigned blocks = (unsigned)ceil_div(threads, 256);
  IMP_DISPATCH_D_MSG(d, (message_agg_kernel<D><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                        g->row_ptr, g->col_src, g->edge_bm, d_h, d_table_il_cat, d_table_il_an, g->n_atoms,
                        g->n_cat_atoms, d_agg, accumulate)));
  IMP_LAUNCH_CHECK();
  return 0;
}

static int edge_messages_grouped(const imp_graph_t* g, const float* d_x, const float* d_table_cat, const float* d_table_an, float* d_msg,
                                 void* d_workspace, int transposed, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * g->bond_vocab;
  int* chunk_ptr = reinterpret_cast<int*>(d_workspace);
  gm_chunk_scan_kernel<<<1, 32, 0, st>>>(g->bucket_ptr, nb, chunk_ptr);
  IMP_LAUNCH_CHECK();
  const unsigned grid = (unsigned)(ceil_div(g->n_unique, GM_CHUNK) + nb);  // upper bound; surplus CTAs exit at once
  if (transposed)
    grouped_msg_f32_kernel<true><<<grid, GM_CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm, g->col_src,
                                                            g->edge_bm, d_x, d_table_cat, d_table_an, d_msg);
  else
    grouped_msg_f32_kernel<false><<<grid, GM_CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm, g->col_src,
                                                             g->edge_bm, d_x, d_table_cat, d_table_an, d_msg);
  IMP_LAUNCH_CHECK();
  return 0;
}

static int edge_messages_any(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_cat, const float* d_table_an,
                             float* d_msg, void* d_workspace, int transposed, void* stream);

extern "C" int imp_edge_messages(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_cat,
                                 const float* d_table_an, float* d_msg, void* stream) {
  return edge_messages_any(g, d_h, d, d_table_cat, d_table_an, d_msg, nullptr, 0, stream);
}
extern "C" int64_t imp_edge_messages_workspace_bytes(int32_t bond_vocab) { return (int64_t)(2 * bond_vocab + 1) * 4; }
extern "C" int imp_edge_messages_grouped(const imp_graph_t* g, const float* d_x, int32_t d, const float* d_table_cat,
                                         const float* d_table_an, int32_t transposed, float* d_msg, void* d_workspace, void* stream) {
  IMP_REQUIRE(d == 32 && d_workspace, IMP_ERR_DIM, "imp_edge_messages_grouped: atom_dim %d (32) and a workspace are required", d);
  return edge_messages_any(g, d_x, d, d_table_cat, d_table_an, d_msg, d_workspace, transposed ? 1 : 0, stream);
}

static int edge_messages_any(const imp_graph_t* g, const float* d_h, int32_t d, const float* d_table_cat, const float* d_table_an,
                             float* d_msg, void* d_workspace, int transposed, void* stream) {
  if (int rc = check_graph(g, "imp_edge_messages")) return rc;
  IMP_REQUIRE(dim_ok_msg(d), IMP_ERR_DIM, "imp_edge_messages: atom_dim %d not in {8,16,32,64,128,256} (fp32 path)", d);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(d_h && d_table_cat && d_table_an && d_msg && g->col_src && g->edge_bm && g->bucket_perm && g->bucket_ptr, IMP_ERR_ARG,
              "imp_edge_messages: null pointer");
  if (d == 32 && d_workspace) return edge_messages_grouped(g, d_h, d_table_cat, d_table_an, d_msg, d_workspace, transposed, stream);
  IMP_REQUIRE(!transposed, IMP_ERR_DIM, "imp_edge_messages_t: atom_dim %d not supported (32, with a workspace)", d);
  const int64_t threads = (int64_t)g->n_unique * d;
  const unsigned blocks = (unsigned)ceil_div(threads, 256);
  IMP_DISPATCH_D_MSG(d, (edge_messages_kernel<D><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                        g->bucket_perm, g->col_src, g->edge_bm, d_h, d_table_cat, d_table_an, g->n_unique,
                        g->bucket_ptr + g->bond_vocab, d_msg)));
  IMP_LAUNCH_CHECK();
  return 0;
}

static int segment_sum_any(const imp_graph_t* g, const float* d_msg, int32_t d, float* d_agg, int accumulate, void* stream);
extern "C" int imp_segment_sum(const imp_graph_t* g, const float* d_msg, int32_t d, float* d_agg, void* stream) {
  return segment_sum_any(g, d_msg, d, d_agg, 0, stream);
}
extern "C" int imp_segment_sum_add(const imp_graph_t* g, const float* d_msg, int32_t d, float* d_agg, void* stream) {
  return segment_sum_any(g, d_msg, d, d_agg, 1, stream);
}
static int segment_sum_any(const imp_graph_t* g, const float* d_msg, int32_t d, float* d_agg, int accumulate, void* stream) {
  if (int rc = check_graph(g, "imp_segment_sum")) return rc;
  IMP_REQUIRE(d > 0 && d % 4 == 0, IMP_ERR_DIM, "imp_segment_sum: atom_dim %d must be a multiple of 4", d);
  if (g->n_atoms == 0) return 0;
  IMP_REQUIRE(d_agg && g->row_ptr && (g->n_unique == 0 || d_msg), IMP_ERR_ARG, "imp_segment_sum: null pointer");
  const int64_t threads = (int64_t)g->n_atoms * (d / 4);
  segment_sum_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(
      g->row_ptr, reinterpret_cast<const float4*>(d_msg), g->n_atoms, d / 4, reinterpret_cast<float4*>(d_agg), accumulate);
  IMP_LAUNCH_CHECK();
  return 0;
}

// K5 for atom_dim 32, register-blocked (the shape of gated_update_bwd_kernel, bwd_fp32.cu): 256 threads per 128-atom tile, a
// thread owns 4 atoms (rows a0 + 4 i of its warp's 16, interleaved: conflict-free float4 row reads) x 4 output columns
// (lane % 8) of every dense product -- 8 LDS.128 per 64 FMA instead of 9 per 32 with one thread per atom -- and the 78 KB of
// weights + staged rows allow two CTAs (16 warps) per SM.  LayerNorm means cross the 8 lanes of a row by three shuffles.
constexpr int K5B_THREADS = 256, K5B_XS = 68, K5B_GS = 36;
struct K5BSmem {
  float Wz[2 * 32 * 32], Wr[2 * 32 * 32], Wh[2 * 32 * 32];
  float bz[32], br[32], bh[32], gamma[32], beta[32];
  float X[K5_TILE * K5B_XS];   // [h | agg]
  float RH[K5_TILE * K5B_GS];  // r * h
};
__device__ __forceinline__ float k5b_comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
template <int LD>
__device__ __forceinline__ void k5b_dense(float (&acc)[4][4], const float* __restrict__ W, const float* __restrict__ xs, int c0) {
#pragma unroll 2
  for (int k4 = 0; k4 < 8; ++k4) {
    float4 xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + i * 4 * LD + 4 * k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(W + (4 * k4 + kk) * 32 + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = k5b_comp(xv[i], kk);
        acc[i][0] = fmaf(x, w.x, acc[i][0]), acc[i][1] = fmaf(x, w.y, acc[i][1]);
        acc[i][2] = fmaf(x, w.z, acc[i][2]), acc[i][3] = fmaf(x, w.w, acc[i][3]);
      }
    }
  }
}
__device__ __forceinline__ float k5b_row_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

__global__ void __launch_bounds__(K5B_THREADS, 2) gated_update32_kernel(const float* __restrict__ h, const float* __restrict__ agg,
                                                                        int n_atoms, int n_cat, int n_cta_cat, imp_gru_weights_t wc,
                                                                        imp_gru_weights_t wa, float eps, float* __restrict__ h_out,
                                                                        float* __restrict__ z_out, float* __restrict__ r_out,
                                                                        float* __restrict__ ht_out) {
  // z_out / r_out / ht_out (all or none): the gates and the candidate of every atom, kept for the backward pass
  // (imp_gated_update_bwd_stored) so that it does not recompute the three Dense layers
  constexpr int D = 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K5BSmem& s = *reinterpret_cast<K5BSmem*>(smem_raw);
  // persistent: CTAs [0, n_cta_cat) walk the cation tiles, the rest the anion tiles; the tower's weights are staged once
  const bool is_cat = (int)blockIdx.x < n_cta_cat;
  const imp_gru_weights_t& w = is_cat ? wc : wa;
  const int base = is_cat ? 0 : n_cat, a_end = is_cat ? n_cat : n_atoms;
  const int n_tiles = (a_end - base + K5_TILE - 1) / K5_TILE;
  const int cta = is_cat ? blockIdx.x : blockIdx.x - n_cta_cat, n_cta = is_cat ? n_cta_cat : gridDim.x - n_cta_cat;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * D * D / 4; i += K5B_THREADS) {
    reinterpret_cast<float4*>(s.Wz)[i] = __ldg(reinterpret_cast<const float4*>(w.Wz) + i);
    reinterpret_cast<float4*>(s.Wr)[i] = __ldg(reinterpret_cast<const float4*>(w.Wr) + i);
    reinterpret_cast<float4*>(s.Wh)[i] = __ldg(reinterpret_cast<const float4*>(w.Wh) + i);
  }
  for (int i = tid; i < D; i += K5B_THREADS)
    s.bz[i] = w.bz[i], s.br[i] = w.br[i], s.bh[i] = w.bh[i], s.gamma[i] = w.gamma[i], s.beta[i] = w.beta[i];
  const int qd = lane >> 3, cg = lane & 7, c0 = 4 * cg;
  __syncthreads();  // weights
  // every row a warp touches below is one of its own 16: the warps of a CTA walk the tiles independently (__syncwarp only)
  for (int tile = cta; tile < n_tiles; tile += n_cta) {
  const int a0 = base + tile * K5_TILE;
  const int rows = min(K5_TILE, a_end - a0);
#pragma unroll
  for (int it = 0; it < 4; ++it) {  // the warp stages its own 16 rows, 8 lanes per 128-byte row
    const int r = 16 * warp + 4 * it + qd;
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), av = hv;
    if (r < rows) {
      hv = __ldg(reinterpret_cast<const float4*>(h + (int64_t)(a0 + r) * D) + cg);
      av = __ldg(reinterpret_cast<const float4*>(agg + (int64_t)(a0 + r) * D) + cg);
    }
    *reinterpret_cast<float4*>(&s.X[r * K5B_XS + c0]) = hv;
    *reinterpret_cast<float4*>(&s.X[r * K5B_XS + D + c0]) = av;
  }
  __syncwarp();
  const int ar0 = 16 * warp + qd;
  const float* Xb = &s.X[ar0 * K5B_XS];
  float* RHb = &s.RH[ar0 * K5B_GS];
  float acc[4][4], zv[4][4], hx[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t4 = *reinterpret_cast<const float4*>(Xb + i * 4 * K5B_XS + c0);
    hx[i][0] = t4.x, hx[i][1] = t4.y, hx[i][2] = t4.z, hx[i][3] = t4.w;
  }
  // r gate -> r * h
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = s.br[c0 + c];
  k5b_dense<K5B_XS>(acc, s.Wr, Xb, c0);
  k5b_dense<K5B_XS>(acc, s.Wr + D * D, Xb + D, c0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 rv = make_float4(sigmoidf_precise(acc[i][0]), sigmoidf_precise(acc[i][1]), sigmoidf_precise(acc[i][2]),
                                  sigmoidf_precise(acc[i][3]));
    *reinterpret_cast<float4*>(RHb + i * 4 * K5B_GS + c0) = make_float4(rv.x * hx[i][0], rv.y * hx[i][1], rv.z * hx[i][2], rv.w * hx[i][3]);
    if (r_out && 16 * warp + qd + 4 * i < rows) reinterpret_cast<float4*>(r_out + (int64_t)(a0 + 16 * warp + qd + 4 * i) * D)[cg] = rv;
  }
  // z gate (independent of r * h: runs while the warp's RH rows settle)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = s.bz[c0 + c];
  k5b_dense<K5B_XS>(acc, s.Wz, Xb, c0);
  k5b_dense<K5B_XS>(acc, s.Wz + D * D, Xb + D, c0);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) zv[i][c] = sigmoidf_precise(acc[i][c]);
  __syncwarp();
  // candidate
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = s.bh[c0 + c];
  k5b_dense<K5B_GS>(acc, s.Wh, RHb, c0);
  k5b_dense<K5B_XS>(acc, s.Wh + D * D, Xb + D, c0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float n[4], mean = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      acc[i][c] = tanhf(acc[i][c]);
      n[c] = (1.0f - zv[i][c]) * hx[i][c] + zv[i][c] * acc[i][c];
      mean += n[c];
    }
    if (z_out && ar0 + 4 * i < rows) {
      reinterpret_cast<float4*>(z_out + (int64_t)(a0 + ar0 + 4 * i) * D)[cg] = make_float4(zv[i][0], zv[i][1], zv[i][2], zv[i][3]);
      reinterpret_cast<float4*>(ht_out + (int64_t)(a0 + ar0 + 4 * i) * D)[cg] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    mean = k5b_row_sum(mean) * (1.0f / D);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      n[c] -= mean;
      var = fmaf(n[c], n[c], var);
    }
    var = k5b_row_sum(var);
    const float inv = 1.0f / sqrtf(var * (1.0f / D) + eps);
    const int r = ar0 + 4 * i;
    if (r < rows)
      reinterpret_cast<float4*>(h_out + (int64_t)(a0 + r) * D)[cg] =
          make_float4(n[0] * inv * s.gamma[c0] + s.beta[c0] + hx[i][0], n[1] * inv * s.gamma[c0 + 1] + s.beta[c0 + 1] + hx[i][1],
                      n[2] * inv * s.gamma[c0 + 2] + s.beta[c0 + 2] + hx[i][2], n[3] * inv * s.gamma[c0 + 3] + s.beta[c0 + 3] + hx[i][3]);
  }
  __syncwarp();  // the warp's rows may be overwritten by its next tile
  }
}

template <int D>
static int launch_k5(const float* h, const float* agg, int n_atoms, int n_cat, const imp_gru_weights_t* wc,
                     const imp_gru_weights_t* wa, float eps, float* out, cudaStream_t st, float* z_out = nullptr,
                     float* r_out = nullptr, float* ht_out = nullptr) {
  const int tiles_cat = (int)ceil_div(n_cat, K5_TILE), tiles_an = (int)ceil_div(n_atoms - n_cat, K5_TILE);
  if (D == 32) {
    IMP_CUDA(cudaFuncSetAttribute(gated_update32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K5BSmem)));
    int dev = 0, sms = 148;
    IMP_CUDA(cudaGetDevice(&dev));
    IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = tiles_cat + tiles_an;
    const int grid = tiles < 2 * sms ? tiles : 2 * sms;  // two CTAs per SM
    int n_cta_cat = tiles > 0 ? (int)((int64_t)grid * tiles_cat / tiles) : 0;
    if (tiles_cat > 0 && n_cta_cat < 1) n_cta_cat = 1;
    if (tiles_an > 0 && n_cta_cat > grid - 1) n_cta_cat = grid - 1;
    if (tiles_an == 0) n_cta_cat = grid;
    gated_update32_kernel<<<grid, K5B_THREADS, sizeof(K5BSmem), st>>>(h, agg, n_atoms, n_cat, n_cta_cat, *wc, *wa, eps, out, z_out,
                                                                      r_out, ht_out);
    IMP_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = sizeof(K5Smem<D>);
  IMP_CUDA(cudaFuncSetAttribute(gated_update_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gated_update_kernel<D><<<tiles_cat + tiles_an, K5_TILE, smem, st>>>(h, agg, n_atoms, n_cat, tiles_cat, *wc, *wa, eps, out);
  IMP_LAUNCH_CHECK();
  return 0;
}

static bool gru_ok(const imp_gru_weights_t* w) {
  return w && w->Wz && w->bz && w->Wr && w->br && w->Wh && w->bh && w->gamma && w->beta;
}

extern "C" int imp_gated_update(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out,
                                void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG, "imp_gated_update: bad sizes");
  IMP_REQUIRE(dim_ok(d), IMP_ERR_DIM, "imp_gated_update: atom_dim %d not in {8,16,32,64} (fp32 path)", d);
  if (n_atoms == 0) return 0;
  IMP_REQUIRE(d_h && d_agg && d_h_out && gru_ok(w_cat) && gru_ok(w_an), IMP_ERR_ARG, "imp_gated_update: null pointer");
  int rc = IMP_ERR_DIM;
  IMP_DISPATCH_D(d, rc = launch_k5<D>(d_h, d_agg, n_atoms, n_cat_atoms, w_cat, w_an, eps, d_h_out, (cudaStream_t)stream));
  return rc;
}

extern "C" int imp_gated_update_train(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                      const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out,
                                      float* d_z, float* d_r, float* d_ht, void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG, "imp_gated_update_train: bad sizes");
  IMP_REQUIRE(d == 32, IMP_ERR_DIM, "imp_gated_update_train: atom_dim %d not supported (32)", d);
  if (n_atoms == 0) return 0;
  IMP_REQUIRE(d_h && d_agg && d_h_out && d_z && d_r && d_ht && gru_ok(w_cat) && gru_ok(w_an), IMP_ERR_ARG,
              "imp_gated_update_train: null pointer");
  return launch_k5<32>(d_h, d_agg, n_atoms, n_cat_atoms, w_cat, w_an, eps, d_h_out, (cudaStream_t)stream, d_z, d_r, d_ht);
}

static int launch_k6(const imp_graph_t* g, const float* d_h, const float* d_pooled, int d, int fp, int mix, int fp2,
                     const imp_readout_weights_t* wc, const imp_readout_weights_t* wa, const float* W1, const float* b1,
                     const float* W2, const float* b2, const float* T, float* out, float* aux, void* stream,
                     const char* who) {
  if (int rc = check_graph(g, who)) return rc;
  IMP_REQUIRE(d > 0 && fp > 0 && mix > 0 && d <= K6_MAXV && fp <= K6_MAXV && mix <= K6_MAXV && fp2 <= K6_MAXV, IMP_ERR_DIM,
              "%s: d/fp/mix/fp2 = %d/%d/%d/%d must be in 1..%d", who, d, fp, mix, fp2, K6_MAXV);
  if (g->n_pairs == 0) return 0;
  IMP_REQUIRE((d_pooled || (d_h && g->mol_ptr && g->atom_id)) && out && wc && wa && wc->W_fp && wc->b_fp && wc->W_mix && wc->b_mix &&
                  wa->W_fp && wa->b_fp && wa->W_mix && wa->b_mix && W1 && b1,
              IMP_ERR_ARG, "%s: null pointer", who);
  K6Args a;
  a.mol_ptr = g->mol_ptr, a.atom_id = g->atom_id, a.h = d_h, a.pooled = d_pooled, a.n_pairs = g->n_pairs;
  a.d = d, a.fp = fp, a.mix = mix, a.fp2 = fp2, a.wc = *wc, a.wa = *wa;
  a.W1 = W1, a.b1 = b1, a.W2 = W2, a.b2 = b2, a.T = T, a.out = out, a.aux = aux;
  const int n_head_out = fp2 > 0 ? fp2 : 3;
  {
    int mv = d > fp ? d : fp;
    mv = mv > mix ? mv : mix;
    mv = mv > fp2 ? mv : fp2;
    a.scratch_stride = (mv + 31) / 32 * 32;
  }
  const size_t smem = sizeof(float) * (2 * d * fp + 2 * fp + 2 * fp * mix + 2 * mix + mix * n_head_out + n_head_out +
                                       (fp2 > 0 ? fp2 : 0) + 2 + 2 * K6_WARPS * 4 * a.scratch_stride);  // scratch rows are doubles
  IMP_CUDA(cudaFuncSetAttribute(pool_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  IMP_REQUIRE(smem <= 200 * 1024, IMP_ERR_DIM, "%s: readout weights need %zu B of shared memory", who, smem);
  if (d_pooled && !aux && d <= 32 && fp <= 32 && mix <= 32 && fp2 <= 32) {  // both reference models: register-resident weights
    IMP_REQUIRE(d % 4 == 0, IMP_ERR_DIM, "%s: atom_dim %d must be a multiple of 4", who, d);
    int nb = (int)ceil_div(g->n_pairs, R32_THREADS);
    if (nb > 148 * 5) nb = 148 * 5;
    IMP_CUDA(cudaFuncSetAttribute(readout32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(R32Smem)));
    readout32_kernel<<<nb, R32_THREADS, sizeof(R32Smem), (cudaStream_t)stream>>>(a);
    IMP_LAUNCH_CHECK();
    return 0;
  }
  int blocks = (int)ceil_div(g->n_pairs, K6_WARPS);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pool_head_kernel<<<blocks, K6_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_global_sum_pool(const int32_t* d_mol_ptr, const int32_t* d_atom_id, int32_t n_mols, const float* d_h,
                                   int32_t d, float* d_out, void* stream) {
  IMP_REQUIRE(n_mols >= 0 && d > 0, IMP_ERR_ARG, "imp_global_sum_pool: bad sizes");
  if (n_mols == 0) return 0;
  IMP_REQUIRE(d_mol_ptr && d_atom_id && d_h && d_out, IMP_ERR_ARG, "imp_global_sum_pool: null pointer");
  global_sum_pool_kernel<<<(unsigned)ceil_div(n_mols, 8), 256, 0, (cudaStream_t)stream>>>(d_mol_ptr, d_atom_id, n_mols, d_h,
                                                                                          d, d_out);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_pool_head_visc(const imp_graph_t* g, const float* d_h, int32_t d, int32_t fp, int32_t mix,
                                  const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an,
                                  const float* d_W_head, const float* d_b_head, const float* d_T, float* d_out,
                                  float* d_aux, void* stream) {
  IMP_REQUIRE(d_T || (g && g->n_pairs == 0), IMP_ERR_ARG, "imp_pool_head_visc: temperature is null");
  return launch_k6(g, d_h, nullptr, d, fp, mix, 0, w_cat, w_an, d_W_head, d_b_head, nullptr, nullptr, d_T, d_out, d_aux, stream,
                   "imp_pool_head_visc");
}

extern "C" int imp_pool_head_mp(const imp_graph_t* g, const float* d_h, int32_t d, int32_t fp, int32_t mix, int32_t fp2,
                                const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an, const float* d_W1,
                                const float* d_b1, const float* d_W2, const float* d_b2, float* d_out, float* d_aux,
                                void* stream) {
  IMP_REQUIRE(fp2 > 0 && d_W2 && d_b2, IMP_ERR_ARG, "imp_pool_head_mp: head weights missing");
  return launch_k6(g, d_h, nullptr, d, fp, mix, fp2, w_cat, w_an, d_W1, d_b1, d_W2, d_b2, nullptr, d_out, d_aux, stream,
                   "imp_pool_head_mp");
}

// Readout on molecule sums that are already pooled (output of imp_mpnn_forward_fused): Dense(fp) + Dense(mix) per
// tower, AddTwoTensors, head.  Same kernel as imp_pool_head_*, with the pooling stage replaced by a row read.
static int readout_graph(int32_t n_pairs, imp_graph_t* g) {
  *g = imp_graph_t{};
  g->n_pairs = n_pairs;
  g->bond_vocab = 1;  // unused by the readout; keeps check_graph's size invariants
  return 0;
}

extern "C" int imp_readout_visc(const float* d_pooled, int32_t n_pairs, int32_t d, int32_t fp, int32_t mix,
                                const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an,
                                const float* d_W_head, const float* d_b_head, const float* d_T, float* d_out, float* d_aux,
                                void* stream) {
  IMP_REQUIRE(n_pairs >= 0, IMP_ERR_ARG, "imp_readout_visc: negative size");
  IMP_REQUIRE((d_T && d_pooled) || n_pairs == 0, IMP_ERR_ARG, "imp_readout_visc: temperature / pooled is null");
  imp_graph_t g;
  readout_graph(n_pairs, &g);
  return launch_k6(&g, nullptr, d_pooled, d, fp, mix, 0, w_cat, w_an, d_W_head, d_b_head, nullptr, nullptr, d_T, d_out, d_aux,
                   stream, "imp_readout_visc");
}

extern "C" int imp_readout_mp(const float* d_pooled, int32_t n_pairs, int32_t d, int32_t fp, int32_t mix, int32_t fp2,
                              const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an, const float* d_W1,
                              const float* d_b1, const float* d_W2, const float* d_b2, float* d_out, float* d_aux,
                              void* stream) {
  IMP_REQUIRE(n_pairs >= 0 && fp2 > 0 && d_W2 && d_b2, IMP_ERR_ARG, "imp_readout_mp: head weights missing");
  IMP_REQUIRE(d_pooled || n_pairs == 0, IMP_ERR_ARG, "imp_readout_mp: pooled is null");
  imp_graph_t g;
  readout_graph(n_pairs, &g);
  return launch_k6(&g, nullptr, d_pooled, d, fp, mix, fp2, w_cat, w_an, d_W1, d_b1, d_W2, d_b2, nullptr, d_out, d_aux, stream,
                   "imp_readout_mp");
}
