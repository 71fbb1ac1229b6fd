// fp32 backward kernels of the MPNN hot path + fused per-variable clip / Adam: the training step of
// train_viscosity.py:227-230,328-338 (loss = mse + l2 kernel regularisers, Adam(1e-3, clipnorm=1.0)).
//
//   B6 readout_bwd        head + Dense(mix) + Dense(fp) backward, loss          train_viscosity.py:189-214, models/layers.py:10-49
//   B5 pool_bwd           GlobalSumPool backward                                models/layers.py:161-164
//   B4 gated_update_bwd   GatedUpdate backward (recomputes the gates)           models/layers.py:142-156
//   B3 message backward   dh += T^T dagg  (imp_message_agg on the transposed table, the live edge set is symmetric)
//      dtable_partial / dtable_reduce / dbond_project   dTable[b] = sum_e mult g[dst_e] (x) h[src_e], then
//                         dW_k = sum_b c[b,k] dTable[b],  dbond_emb[b,k] += <dTable[b], W_k>      models/layers.py:100-117
//   B1 embed_bwd          Embedding(atom) backward                              train_viscosity.py:163,171
//   sumsq / clip_adam     per-variable clip_by_norm + Adam                      train_viscosity.py:227-230 [Keras semantics]
//
// Every reduction over atoms / pairs / entries is two-stage with a fixed order (per-CTA partials, then one thread per
// output element sums the partials in index order): results are bit-reproducible run to run, unlike the reference's
// scatter-based gradients on a GPU.
#include <math.h>

#include "common.cuh"

namespace imp {

__device__ __forceinline__ float bw_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float bw_softplus(float x) { return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x)); }

// ================================================================================================ B5
__global__ void pool_bwd_kernel(const int* __restrict__ mol_ptr, const int* __restrict__ atom_id, int n_mols,
                                const float* __restrict__ d_pooled, int d, float* __restrict__ dh) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= n_mols) return;
  const int v0 = mol_ptr[m], v1 = mol_ptr[m + 1];
  for (int j = lane; j < d; j += 32) {
    const float g = d_pooled[(int64_t)m * d + j];
    for (int v = v0; v < v1; ++v) dh[(int64_t)v * d + j] = atom_id[v] > 0 ? g : 0.f;
  }
}

// ================================================================================================ B6
constexpr int RB_WARPS = 8;
constexpr int RB_MAXV = 32;  // d, fp, mix, fp2 <= 32 in the backward readout

struct ReadoutBwdArgs {
  const float* pooled;  // [2P, d]
  const float* T;       // [P] or null (melting point)
  const float* y;       // [P] targets
  imp_readout_weights_t wc, wa;
  const float *W1, *b1, *W2, *b2;
  int n_pairs, d, fp, mix, fp2;
  float scale;          // 2 / global batch
  float* d_pooled;      // [2P, d]
  float* out;           // [P] predictions (optional)
  float* partial;       // [n_warps_total][readout_grad_floats + 1]
};

__host__ __device__ inline int readout_grad_floats(int d, int fp, int mix, int fp2) {
  const int nh = fp2 > 0 ? fp2 : 3;
  return 2 * (d * fp + fp + fp * mix + mix) + mix * nh + nh + (fp2 > 0 ? fp2 + 1 : 0);
}

__global__ void __launch_bounds__(RB_WARPS * 32) readout_bwd_kernel(ReadoutBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int d = a.d, fp = a.fp, mix = a.mix, fp2 = a.fp2, nh = fp2 > 0 ? fp2 : 3;
  // weights (row-major as given) + transposed copies for the input-gradient products
  float* p = sm;
  float *Wfp[2], *WfpT[2], *bfp[2], *Wmx[2], *WmxT[2], *bmx[2];
  for (int t = 0; t < 2; ++t) {
    Wfp[t] = p, p += d * fp;
    WfpT[t] = p, p += d * fp;
    bfp[t] = p, p += fp;
    Wmx[t] = p, p += fp * mix;
    WmxT[t] = p, p += fp * mix;
    bmx[t] = p, p += mix;
  }
  float* W1 = p;
  p += mix * nh;
  float* W1T = p;
  p += mix * nh;
  float* b1 = p;
  p += nh;
  float* W2 = p;
  p += (fp2 > 0 ? fp2 : 0);
  float* scratch = p + (threadIdx.x / 32) * (8 * RB_MAXV);
  for (int t = 0; t < 2; ++t) {
    const imp_readout_weights_t& w = t == 0 ? a.wc : a.wa;
    for (int i = threadIdx.x; i < d * fp; i += blockDim.x) {
      const float v = w.W_fp[i];
      Wfp[t][i] = v;
      WfpT[t][(i % fp) * d + i / fp] = v;
    }
    for (int i = threadIdx.x; i < fp; i += blockDim.x) bfp[t][i] = w.b_fp[i];
    for (int i = threadIdx.x; i < fp * mix; i += blockDim.x) {
      const float v = w.W_mix[i];
      Wmx[t][i] = v;
      WmxT[t][(i % mix) * fp + i / mix] = v;
    }
    for (int i = threadIdx.x; i < mix; i += blockDim.x) bmx[t][i] = w.b_mix[i];
  }
  for (int i = threadIdx.x; i < mix * nh; i += blockDim.x) {
    const float v = a.W1[i];
    W1[i] = v;
    W1T[(i % nh) * mix + i / nh] = v;
  }
  for (int i = threadIdx.x; i < nh; i += blockDim.x) b1[i] = a.b1[i];
  if (fp2 > 0)
    for (int i = threadIdx.x; i < fp2; i += blockDim.x) W2[i] = a.W2[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * RB_WARPS + (threadIdx.x >> 5), n_warps = gridDim.x * RB_WARPS;
  float* pool = scratch;                  // [2][32]
  float* v1 = scratch + 2 * RB_MAXV;      // [2][32]
  float* v2 = scratch + 4 * RB_MAXV;      // [2][32]
  float* mixed = scratch + 6 * RB_MAXV;   // [32]
  float* tmp = scratch + 7 * RB_MAXV;     // [32]
  // per-lane accumulators: lane j owns column j of every weight-gradient matrix
  float gWfp[2][RB_MAXV], gWmx[2][RB_MAXV], gW1[RB_MAXV];
  float gbfp[2] = {0.f, 0.f}, gbmx[2] = {0.f, 0.f}, gb1 = 0.f, gW2 = 0.f, gb2 = 0.f, sse = 0.f;
#pragma unroll
  for (int i = 0; i < RB_MAXV; ++i) gWfp[0][i] = gWfp[1][i] = gWmx[0][i] = gWmx[1][i] = gW1[i] = 0.f;

  for (int pair = gw; pair < a.n_pairs; pair += n_warps) {
    // ---- forward (as K6)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int m = t * a.n_pairs + pair;
      pool[t * RB_MAXV + lane] = lane < d ? a.pooled[(int64_t)m * d + lane] : 0.f;
      __syncwarp();
      float acc = lane < fp ? bfp[t][lane] : 0.f;
      if (lane < fp)
        for (int k = 0; k < d; ++k) acc = fmaf(pool[t * RB_MAXV + k], Wfp[t][k * fp + lane], acc);
      v1[t * RB_MAXV + lane] = lane < fp ? fmaxf(acc, 0.f) : 0.f;
      __syncwarp();
      acc = lane < mix ? bmx[t][lane] : 0.f;
      if (lane < mix)
        for (int k = 0; k < fp; ++k) acc = fmaf(v1[t * RB_MAXV + k], Wmx[t][k * mix + lane], acc);
      v2[t * RB_MAXV + lane] = lane < mix ? fmaxf(acc, 0.f) : 0.f;
      __syncwarp();
    }
    mixed[lane] = v2[lane] + v2[RB_MAXV + lane];
    __syncwarp();
    float hp = lane < nh ? b1[lane] : 0.f;  // head pre-activation of output lane
    if (lane < nh)
      for (int k = 0; k < mix; ++k) hp = fmaf(mixed[k], W1[k * nh + lane], hp);
    float out, dhead;  // dhead = dL/d(head pre-activation of this lane)
    const float yv = a.y[pair];
    if (fp2 == 0) {
      const float p0 = __shfl_sync(0xffffffffu, hp, 0), p1 = __shfl_sync(0xffffffffu, hp, 1), p2 = __shfl_sync(0xffffffffu, hp, 2);
      const float sB = bw_softplus(p1), sC = bw_softplus(p2);
      const float B = fminf(fmaxf(sB, 0.0f), 20.0f), Cc = fminf(fmaxf(sC, 0.1f), 50.0f);
      const float den = a.T[pair] / 100.0f + Cc + 1e-6f;
      out = p0 + B / den;
      const float dout = a.scale * (out - yv);
      const float dB = (sB >= 0.0f && sB <= 20.0f) ? dout / den : 0.f;
      const float dC = (sC >= 0.1f && sC <= 50.0f) ? -dout * B / (den * den) : 0.f;
      dhead = lane == 0 ? dout : lane == 1 ? dB * bw_sigmoid(p1) : lane == 2 ? dC * bw_sigmoid(p2) : 0.f;
    } else {
      const float hid = lane < fp2 ? fmaxf(hp, 0.f) : 0.f;
      float part = lane < fp2 ? hid * W2[lane] : 0.f;
      float tot = 0.f;
      for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, part, l);  // fixed order
      out = tot + a.b2[0];
      const float dout = a.scale * (out - yv);
      gW2 = fmaf(hid, dout, gW2);
      gb2 += dout;
      dhead = (lane < fp2 && hp > 0.f) ? W2[lane] * dout : 0.f;
    }
    if (a.out && lane == 0) a.out[pair] = out;
    if (lane == 0) sse = fmaf(out - yv, out - yv, sse);
    // ---- head backward
    gb1 += dhead;
#pragma unroll
    for (int i = 0; i < RB_MAXV; ++i) gW1[i] = fmaf(i < mix ? mixed[i] : 0.f, dhead, gW1[i]);
    tmp[lane] = dhead;
    __syncwarp();
    float dmixed = 0.f;
    if (lane < mix)
      for (int k = 0; k < nh; ++k) dmixed = fmaf(W1T[k * mix + lane], tmp[k], dmixed);
    __syncwarp();
    // ---- towers
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int m = t * a.n_pairs + pair;
      const float dv2 = (lane < mix && v2[t * RB_MAXV + lane] > 0.f) ? dmixed : 0.f;
      gbmx[t] += dv2;
#pragma unroll
      for (int i = 0; i < RB_MAXV; ++i) gWmx[t][i] = fmaf(v1[t * RB_MAXV + i], dv2, gWmx[t][i]);
      tmp[lane] = dv2;
      __syncwarp();
      float dv1 = 0.f;
      if (lane < fp) {
        for (int k = 0; k < mix; ++k) dv1 = fmaf(WmxT[t][k * fp + lane], tmp[k], dv1);
        if (!(v1[t * RB_MAXV + lane] > 0.f)) dv1 = 0.f;
      }
      __syncwarp();
      gbfp[t] += dv1;
#pragma unroll
      for (int i = 0; i < RB_MAXV; ++i) gWfp[t][i] = fmaf(pool[t * RB_MAXV + i], dv1, gWfp[t][i]);
      tmp[lane] = dv1;
      __syncwarp();
      float dp = 0.f;
      if (lane < d) {
        for (int k = 0; k < fp; ++k) dp = fmaf(WfpT[t][k * d + lane], tmp[k], dp);
        a.d_pooled[(int64_t)m * d + lane] = dp;
      }
      __syncwarp();
    }
  }
  // ---- per-warp partials (layout = readout_grad_floats order, then the squared-error sum)
  const int stride = readout_grad_floats(d, fp, mix, fp2) + 1;
  float* o = a.partial + (int64_t)gw * stride;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
#pragma unroll
    for (int i = 0; i < RB_MAXV; ++i)
      if (i < d && lane < fp) o[i * fp + lane] = gWfp[t][i];
    o += d * fp;
    if (lane < fp) o[lane] = gbfp[t];
    o += fp;
#pragma unroll
    for (int i = 0; i < RB_MAXV; ++i)
      if (i < fp && lane < mix) o[i * mix + lane] = gWmx[t][i];
    o += fp * mix;
    if (lane < mix) o[lane] = gbmx[t];
    o += mix;
  }
#pragma unroll
  for (int i = 0; i < RB_MAXV; ++i)
    if (i < mix && lane < nh) o[i * nh + lane] = gW1[i];
  o += mix * nh;
  if (lane < nh) o[lane] = gb1;
  o += nh;
  if (fp2 > 0) {
    if (lane < fp2) o[lane] = gW2;
    o += fp2;
    if (lane == 0) o[0] = gb2;
    o += 1;
  }
  if (lane == 0) o[0] = sse;
}

// sums `n_parts` partial vectors of `n` floats in index order: out[i] (+)= sum_p partial[p][i]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_parts, int64_t stride, int n,
                                       float* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += partial[(int64_t)p * stride + i];
  out[i] = accumulate ? out[i] + s : s;
}

// ================================================================================================ B4
constexpr int GB_TILE = 128;
constexpr int GB_XS = 68;  // row stride of [h | agg] (floats; 16-byte aligned, conflict-free float4 row writes)
constexpr int GB_GS = 36;  // row stride of the 32-wide per-atom vectors

template <int D>
struct GruBwdSmem {
  float Wz[2 * D * D], Wr[2 * D * D], Wh[2 * D * D];
  float WzT[2 * D * D], WrT[2 * D * D], WhT[2 * D * D];  // WT[j][k] = W[k][j], rows of 2D floats (transposed products)
  float bz[D], br[D], bh[D], gamma[D], beta[D];
  float X[GB_TILE * GB_XS];   // [h | agg]
  float RH[GB_TILE * GB_GS];  // r * h
  float Gz[GB_TILE * GB_GS];  // z, then dL/dzpre
  float Gr[GB_TILE * GB_GS];  // r, then dL/drpre
  float Gh[GB_TILE * GB_GS];  // dL/dhpre
  float GX[GB_TILE * GB_GS];  // g_out * xhat (for dgamma)
};

template <int D>
__host__ __device__ constexpr int gru_grad_floats() { return 3 * 2 * D * D + 5 * D; }
// layout (= the layer's variable order in the flat parameter buffer): dWz dbz dWr dbr dWh dbh dgamma dbeta

// 256 threads for a 128-atom tile; a thread owns a 4 x 4 register block of every row-by-weight product: 4 atoms
// (rows a0 + 4 i of its warp's 16 atoms, interleaved so that the four lane groups of a warp read four consecutive shared
// rows: conflict-free float4 reads) x 4 output columns (lane % 8).  Per 4 steps of a contraction that is 4 activation float4
// + 4 weight float4 for 64 FMA (the one-thread-per-atom form read 9 LDS per 32 FMA and left the SM with 4 warps).
// Transposed products (gate gradients back through Wz, Wr, Wh) read transposed weight copies staged once per CTA, so they
// have the same shape.  Row-wide terms (LayerNorm means) cross the 8 lanes of a row by three shuffles; everything a
// thread reads from another lane's columns goes through the warp's own shared rows (__syncwarp only).
constexpr int GB_THREADS = 2 * GB_TILE;

__device__ __forceinline__ float gb_comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// acc[i][c] += sum_{k < 32} x[atom i][k] * W[k][c0 + c]      (x rows: xs + i * 4 * LD; W rows of WLD floats)
template <int LD, int WLD>
__device__ __forceinline__ void gb4_dense(float (&acc)[4][4], const float* __restrict__ W, const float* __restrict__ xs, int c0) {
#pragma unroll 2
  for (int k4 = 0; k4 < 8; ++k4) {
    float4 xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + i * 4 * LD + 4 * k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(W + (4 * k4 + kk) * WLD + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = gb_comp(xv[i], kk);
        acc[i][0] = fmaf(x, w.x, acc[i][0]), acc[i][1] = fmaf(x, w.y, acc[i][1]);
        acc[i][2] = fmaf(x, w.z, acc[i][2]), acc[i][3] = fmaf(x, w.w, acc[i][3]);
      }
    }
  }
}
// a1[i][c] += sum_j G[atom i][j] * WT[j][c0 + c];  a2[i][c] += sum_j G[atom i][j] * WT[j][32 + c0 + c]
template <int LD>
__device__ __forceinline__ void gb4_dense_t(float (&a1)[4][4], float (&a2)[4][4], const float* __restrict__ WT,
                                            const float* __restrict__ gs, int c0) {
#pragma unroll 2
  for (int j4 = 0; j4 < 8; ++j4) {
    float4 gv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gv[i] = *reinterpret_cast<const float4*>(gs + i * 4 * LD + 4 * j4);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float4 w1 = *reinterpret_cast<const float4*>(WT + (4 * j4 + jj) * 64 + c0);
      const float4 w2 = *reinterpret_cast<const float4*>(WT + (4 * j4 + jj) * 64 + 32 + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float g = gb_comp(gv[i], jj);
        a1[i][0] = fmaf(g, w1.x, a1[i][0]), a1[i][1] = fmaf(g, w1.y, a1[i][1]);
        a1[i][2] = fmaf(g, w1.z, a1[i][2]), a1[i][3] = fmaf(g, w1.w, a1[i][3]);
        a2[i][0] = fmaf(g, w2.x, a2[i][0]), a2[i][1] = fmaf(g, w2.y, a2[i][1]);
        a2[i][2] = fmaf(g, w2.z, a2[i][2]), a2[i][3] = fmaf(g, w2.w, a2[i][3]);
      }
    }
  }
}
__device__ __forceinline__ float gb_row_sum(float v) {  // over the 8 lanes (lane % 8) that share a row
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// Persistent CTAs; CTAs [0, n_cta_cat) walk the cation tiles, the rest the anion tiles.  Per-CTA partial weight
// gradients are written to partial[cta][gru_grad_floats]; imp_gated_update_bwd reduces them per tower in CTA order.
// STORED: the gates z, r and the candidate tanh(.) were kept by the forward (imp_gated_update_train) and are read instead of
// recomputed -- a third of the kernel's FMAs.
template <int D, bool STORED>
__global__ void __launch_bounds__(GB_THREADS) gated_update_bwd_kernel(const float* __restrict__ h, const float* __restrict__ agg,
                                                                      const float* __restrict__ g_out, int n_atoms, int n_cat,
                                                                      int n_cta_cat, imp_gru_weights_t wc, imp_gru_weights_t wa,
                                                                      float eps, float* __restrict__ dh, float* __restrict__ dagg,
                                                                      float* __restrict__ partial, const float* __restrict__ zs,
                                                                      const float* __restrict__ rs, const float* __restrict__ hts) {
  static_assert(D == 32, "gated_update_bwd is instantiated for atom_dim 32");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GruBwdSmem<D>& s = *reinterpret_cast<GruBwdSmem<D>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_cat = (int)blockIdx.x < n_cta_cat;
  const imp_gru_weights_t& w = is_cat ? wc : wa;
  const int base = is_cat ? 0 : n_cat, a_end = is_cat ? n_cat : n_atoms;
  const int n_tiles = (a_end - base + GB_TILE - 1) / GB_TILE;
  const int cta = is_cat ? blockIdx.x : blockIdx.x - n_cta_cat, n_cta = is_cat ? n_cta_cat : gridDim.x - n_cta_cat;
  for (int i = tid; i < 2 * D * D / 4; i += GB_THREADS) {
    reinterpret_cast<float4*>(s.Wz)[i] = __ldg(reinterpret_cast<const float4*>(w.Wz) + i);
    reinterpret_cast<float4*>(s.Wr)[i] = __ldg(reinterpret_cast<const float4*>(w.Wr) + i);
    reinterpret_cast<float4*>(s.Wh)[i] = __ldg(reinterpret_cast<const float4*>(w.Wh) + i);
  }
  for (int i = tid; i < 2 * D * D; i += GB_THREADS) {  // WT[j][k] = W[k][j]
    const int j = i / (2 * D), k = i % (2 * D);
    s.WzT[i] = __ldg(w.Wz + k * D + j), s.WrT[i] = __ldg(w.Wr + k * D + j), s.WhT[i] = __ldg(w.Wh + k * D + j);
  }
  for (int i = tid; i < D; i += GB_THREADS)
    s.bz[i] = w.bz[i], s.br[i] = w.br[i], s.bh[i] = w.bh[i], s.gamma[i] = w.gamma[i], s.beta[i] = w.beta[i];
  __syncthreads();

  // weight-gradient accumulators: two atom groups (even / odd atoms of a tile) of 128 threads; in a group, thread
  // (k2 = t % 32, jb = t / 32) owns rows 2 k2, 2 k2 + 1 and columns [8 jb, 8 jb + 8) of the three 64 x 32 gradients:
  // 8 shared-memory reads per 48 FMA (the one-row form read 8 per 24).  The groups are combined once, at the end.
  const int wag = tid >> 7, wk2 = tid & 31, wjb = (tid & 127) >> 5;
  float aWz[2][8], aWr[2][8], aWh[2][8];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int i = 0; i < 8; ++i) aWz[kk][i] = aWr[kk][i] = aWh[kk][i] = 0.f;
  // vector-gradient accumulators: thread (j = tid % 32, q = tid / 32) sums atoms a = q, q+8, ...
  const int vj = tid % D, vq = tid / D;
  float abz = 0.f, abr = 0.f, abh = 0.f, agam = 0.f, abet = 0.f;

  const int qd = lane >> 3, cg = lane & 7, c0 = 4 * cg;
  const int ar0 = 16 * warp + qd;  // tile row of this thread's atom 0; atom i is row ar0 + 4 i
  const float* Xb = &s.X[ar0 * GB_XS];
  float* RHb = &s.RH[ar0 * GB_GS];
  float* Gzb = &s.Gz[ar0 * GB_GS];
  float* Grb = &s.Gr[ar0 * GB_GS];
  float* Ghb = &s.Gh[ar0 * GB_GS];
  float* GXb = &s.GX[ar0 * GB_GS];

  for (int tile = cta; tile < n_tiles; tile += n_cta) {
    const int a0 = base + tile * GB_TILE;
    const int rows = min(GB_TILE, a_end - a0);
    // ---- the warp stages its own 16 rows of [h | agg] (8 lanes per 128-byte row); rows beyond the tile end are zeros and
    // contribute exact zeros to every gradient
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = 16 * warp + 4 * it + qd;
      float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), av = hv;
      if (r < rows) {
        hv = __ldg(reinterpret_cast<const float4*>(h + (int64_t)(a0 + r) * D) + cg);
        av = __ldg(reinterpret_cast<const float4*>(agg + (int64_t)(a0 + r) * D) + cg);
      }
      *reinterpret_cast<float4*>(&s.X[r * GB_XS + c0]) = hv;
      *reinterpret_cast<float4*>(&s.X[r * GB_XS + D + c0]) = av;
    }
    float go[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ar0 + 4 * i;
      const float4 gv = r < rows ? __ldg(reinterpret_cast<const float4*>(g_out + (int64_t)(a0 + r) * D) + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
      go[i][0] = gv.x, go[i][1] = gv.y, go[i][2] = gv.z, go[i][3] = gv.w;
    }
    __syncwarp();
    float acc[4][4], zv[4][4], rv[4][4], hx[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t4 = *reinterpret_cast<const float4*>(Xb + i * 4 * GB_XS + c0);
      hx[i][0] = t4.x, hx[i][1] = t4.y, hx[i][2] = t4.z, hx[i][3] = t4.w;
    }
    if constexpr (STORED) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ar0 + 4 * i;
        float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = z4, t4 = z4;
        if (r < rows) {
          z4 = __ldg(reinterpret_cast<const float4*>(zs + (int64_t)(a0 + r) * D) + cg);
          r4 = __ldg(reinterpret_cast<const float4*>(rs + (int64_t)(a0 + r) * D) + cg);
          t4 = __ldg(reinterpret_cast<const float4*>(hts + (int64_t)(a0 + r) * D) + cg);
        }
        zv[i][0] = z4.x, zv[i][1] = z4.y, zv[i][2] = z4.z, zv[i][3] = z4.w;
        rv[i][0] = r4.x, rv[i][1] = r4.y, rv[i][2] = r4.z, rv[i][3] = r4.w;
        acc[i][0] = t4.x, acc[i][1] = t4.y, acc[i][2] = t4.z, acc[i][3] = t4.w;
        *reinterpret_cast<float4*>(RHb + i * 4 * GB_GS + c0) =
            make_float4(rv[i][0] * hx[i][0], rv[i][1] * hx[i][1], rv[i][2] * hx[i][2], rv[i][3] * hx[i][3]);
      }
    } else {
    // z
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = s.bz[c0 + c];
      gb4_dense<GB_XS, D>(acc, s.Wz, Xb, c0);
      gb4_dense<GB_XS, D>(acc, s.Wz + D * D, Xb + D, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) zv[i][c] = bw_sigmoid(acc[i][c]);
      // r, r*h
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = s.br[c0 + c];
      gb4_dense<GB_XS, D>(acc, s.Wr, Xb, c0);
      gb4_dense<GB_XS, D>(acc, s.Wr + D * D, Xb + D, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) rv[i][c] = bw_sigmoid(acc[i][c]);
        *reinterpret_cast<float4*>(RHb + i * 4 * GB_GS + c0) =
            make_float4(rv[i][0] * hx[i][0], rv[i][1] * hx[i][1], rv[i][2] * hx[i][2], rv[i][3] * hx[i][3]);
      }
      __syncwarp();
      // candidate
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = s.bh[c0 + c];
      gb4_dense<GB_GS, D>(acc, s.Wh, RHb, c0);
      gb4_dense<GB_XS, D>(acc, s.Wh + D * D, Xb + D, c0);
    }
    // blend, LayerNorm forward + backward, gate gradients (per atom row i)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float nrm[4], mean = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if constexpr (!STORED) acc[i][c] = tanhf(acc[i][c]);  // ht
        nrm[c] = fmaf(zv[i][c], acc[i][c] - hx[i][c], hx[i][c]);
        mean += nrm[c];
      }
      mean = gb_row_sum(mean) * (1.0f / D);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        nrm[c] -= mean;
        var = fmaf(nrm[c], nrm[c], var);
      }
      var = gb_row_sum(var);
      const float inv = 1.0f / sqrtf(var * (1.0f / D) + eps);
      // LayerNorm backward: dn = inv * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
      float m1 = 0.f, m2 = 0.f, gx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        nrm[c] *= inv;  // xhat
        gx[c] = go[i][c] * nrm[c];
        const float dx = go[i][c] * s.gamma[c0 + c];
        m1 += dx;
        m2 = fmaf(dx, nrm[c], m2);
      }
      *reinterpret_cast<float4*>(GXb + i * 4 * GB_GS + c0) = make_float4(gx[0], gx[1], gx[2], gx[3]);
      m1 = gb_row_sum(m1) * (1.0f / D), m2 = gb_row_sum(m2) * (1.0f / D);
      float gz[4], gh[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float dn = inv * (go[i][c] * s.gamma[c0 + c] - m1 - nrm[c] * m2);
        const float z = zv[i][c], ht = acc[i][c], hj = hx[i][c];
        go[i][c] = fmaf(dn, 1.0f - z, go[i][c]);   // dh starts as the residual path + the (1 - z) path
        gz[c] = dn * (ht - hj) * z * (1.0f - z);    // dL/dzpre
        gh[c] = dn * z * (1.0f - ht * ht);          // dL/dhpre
      }
      *reinterpret_cast<float4*>(Gzb + i * 4 * GB_GS + c0) = make_float4(gz[0], gz[1], gz[2], gz[3]);
      *reinterpret_cast<float4*>(Ghb + i * 4 * GB_GS + c0) = make_float4(gh[0], gh[1], gh[2], gh[3]);
    }
    __syncwarp();
    // through Wh: d(r*h) (columns k < 32) and dagg (k >= 32)
    float dag[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = dag[i][c] = 0.f;
    gb4_dense_t<GB_GS>(acc, dag, s.WhT, Ghb, c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float gr[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float drh = acc[i][c], r = rv[i][c];
        go[i][c] = fmaf(drh, r, go[i][c]);
        gr[c] = drh * hx[i][c] * r * (1.0f - r);  // dL/drpre
      }
      *reinterpret_cast<float4*>(Grb + i * 4 * GB_GS + c0) = make_float4(gr[0], gr[1], gr[2], gr[3]);
    }
    __syncwarp();
    // through Wz, Wr
    gb4_dense_t<GB_GS>(go, dag, s.WzT, Gzb, c0);
    gb4_dense_t<GB_GS>(go, dag, s.WrT, Grb, c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ar0 + 4 * i;
      if (r < rows) {
        reinterpret_cast<float4*>(dh + (int64_t)(a0 + r) * D)[cg] = make_float4(go[i][0], go[i][1], go[i][2], go[i][3]);
        reinterpret_cast<float4*>(dagg + (int64_t)(a0 + r) * D)[cg] = make_float4(dag[i][0], dag[i][1], dag[i][2], dag[i][3]);
      }
    }
    __syncthreads();
    // ---- weight gradients of this tile: dW[k][j] += sum_a X[a][k] G[a][j]
    for (int a = wag; a < rows; a += 2) {
      const float2 x2 = *reinterpret_cast<const float2*>(&s.X[a * GB_XS + 2 * wk2]);            // [h | agg][2 k2 ..]
      const float2 h2 = wk2 < D / 2 ? *reinterpret_cast<const float2*>(&s.RH[a * GB_GS + 2 * wk2]) : x2;  // [r*h | agg]
      const float xk[2] = {x2.x, x2.y}, xh[2] = {h2.x, h2.y};
      const float4* gz = reinterpret_cast<const float4*>(&s.Gz[a * GB_GS + 8 * wjb]);
      const float4* gr = reinterpret_cast<const float4*>(&s.Gr[a * GB_GS + 8 * wjb]);
      const float4* gh = reinterpret_cast<const float4*>(&s.Gh[a * GB_GS + 8 * wjb]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 z4 = gz[c], r4 = gr[c], h4 = gh[c];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          aWz[kk][4 * c] = fmaf(xk[kk], z4.x, aWz[kk][4 * c]), aWz[kk][4 * c + 1] = fmaf(xk[kk], z4.y, aWz[kk][4 * c + 1]);
          aWz[kk][4 * c + 2] = fmaf(xk[kk], z4.z, aWz[kk][4 * c + 2]), aWz[kk][4 * c + 3] = fmaf(xk[kk], z4.w, aWz[kk][4 * c + 3]);
          aWr[kk][4 * c] = fmaf(xk[kk], r4.x, aWr[kk][4 * c]), aWr[kk][4 * c + 1] = fmaf(xk[kk], r4.y, aWr[kk][4 * c + 1]);
          aWr[kk][4 * c + 2] = fmaf(xk[kk], r4.z, aWr[kk][4 * c + 2]), aWr[kk][4 * c + 3] = fmaf(xk[kk], r4.w, aWr[kk][4 * c + 3]);
          aWh[kk][4 * c] = fmaf(xh[kk], h4.x, aWh[kk][4 * c]), aWh[kk][4 * c + 1] = fmaf(xh[kk], h4.y, aWh[kk][4 * c + 1]);
          aWh[kk][4 * c + 2] = fmaf(xh[kk], h4.z, aWh[kk][4 * c + 2]), aWh[kk][4 * c + 3] = fmaf(xh[kk], h4.w, aWh[kk][4 * c + 3]);
        }
      }
    }
    for (int a = vq; a < rows; a += GB_THREADS / D) {
      abz += s.Gz[a * GB_GS + vj];
      abr += s.Gr[a * GB_GS + vj];
      abh += s.Gh[a * GB_GS + vj];
      agam += s.GX[a * GB_GS + vj];
      abet += __ldg(g_out + (int64_t)(a0 + a) * D + vj);
    }
    __syncthreads();
  }
  // ---- per-CTA partials.  Vector gradients: combine the 8 atom phases through shared memory in phase order.
  float* o = partial + (int64_t)blockIdx.x * gru_grad_floats<D>();
  {  // odd-atom group -> shared memory (the staged rows are dead), even-atom group adds it and writes: fixed order
    float* xg = s.X;  // [128 threads][48]
    if (wag == 1) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xg[(tid & 127) * 48 + kk * 8 + i] = aWz[kk][i];
          xg[(tid & 127) * 48 + 16 + kk * 8 + i] = aWr[kk][i];
          xg[(tid & 127) * 48 + 32 + kk * 8 + i] = aWh[kk][i];
        }
    }
    __syncthreads();
    if (wag == 0) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = 2 * wk2 + kk;
          o[k * D + 8 * wjb + i] = aWz[kk][i] + xg[tid * 48 + kk * 8 + i];
          o[(2 * D * D + D) + k * D + 8 * wjb + i] = aWr[kk][i] + xg[tid * 48 + 16 + kk * 8 + i];
          o[2 * (2 * D * D + D) + k * D + 8 * wjb + i] = aWh[kk][i] + xg[tid * 48 + 32 + kk * 8 + i];
        }
    }
    __syncthreads();
  }
  constexpr int NQ = GB_THREADS / D;
  float* red = s.X;  // reuse: [5][NQ][32]
  red[(0 * NQ + vq) * D + vj] = abz, red[(1 * NQ + vq) * D + vj] = abr, red[(2 * NQ + vq) * D + vj] = abh;
  red[(3 * NQ + vq) * D + vj] = agam, red[(4 * NQ + vq) * D + vj] = abet;
  __syncthreads();
  for (int i = tid; i < 5 * D; i += GB_THREADS) {
    const int v = i / D, jj = i % D;
    const int off = v < 3 ? v * (2 * D * D + D) + 2 * D * D : 3 * (2 * D * D + D) + (v - 3) * D;  // bz, br, bh | gamma, beta
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) sum += red[(v * NQ + q) * D + jj];
    o[off + jj] = sum;
  }
}

// ================================================================================================ B3 (dTable)
constexpr int DT_CHUNK_THREADS = 256;
constexpr int DT_SUB = 128;  // entries staged per pass
constexpr int DT_LD = 36;    // padded row (floats): 16-byte aligned rows, conflict-free float4 reads of one row
// One CTA per chunk of a (tower, bond) bucket: partial[chunk][l][m] = sum_e mult_e g[dst_e][l] h[src_e][m].
// The chunk is walked in passes of 128 entries: indices, then both 128-byte rows of every entry are staged in shared memory
// (8 lanes per row: whole lines), mult folded into the g row.  Four groups of 64 threads each take a quarter of the pass;
// a thread owns a 4 x 4 block of the 32 x 32 outer-product sum, i.e. two LDS.128 feed 16 FMA per entry (the previous form
// did one global scalar + one global float4 load and four redundant index loads per 4 FMA in every one of 256 threads).
// The four group sums are combined in group order at the end: bit-reproducible.
template <int D>
__global__ void __launch_bounds__(DT_CHUNK_THREADS) dtable_partial_kernel(const int* __restrict__ chunk_begin,
                                                                          const int* __restrict__ chunk_end,
                                                                          const int* __restrict__ bucket_perm,
                                                                          const int* __restrict__ entry_dst,
                                                                          const int* __restrict__ col_src,
                                                                          const int* __restrict__ edge_bm,
                                                                          const float* __restrict__ g, const float* __restrict__ h,
                                                                          float* __restrict__ partial) {
  static_assert(D == 32, "dtable kernel is instantiated for atom_dim 32");
  __shared__ __align__(16) float sG[DT_SUB * DT_LD];
  __shared__ __align__(16) float sH[DT_SUB * DT_LD];
  __shared__ int sDst[DT_SUB], sSrc[DT_SUB];
  __shared__ float sMult[DT_SUB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 6, tg = tid & 63, lq = tg >> 3, mq = tg & 7;
  const int lg = lane >> 3, q = lane & 7;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int e1 = chunk_end[blockIdx.x];
  for (int i0 = chunk_begin[blockIdx.x]; i0 < e1; i0 += DT_SUB) {
    const int n = min(DT_SUB, e1 - i0);
    if (tid < DT_SUB) {
      int dst = -1, src = 0;
      float mult = 0.f;
      if (tid < n) {
        const int e = __ldg(bucket_perm + i0 + tid);
        dst = __ldg(entry_dst + e), src = __ldg(col_src + e);
        mult = (float)((unsigned)__ldg(edge_bm + e) >> 16);
      }
      sDst[tid] = dst, sSrc[tid] = src, sMult[tid] = mult;
    }
    __syncthreads();
    // warp w stages entries [16 w, 16 w + 16): 4 lane groups x 4 iterations, lane % 8 = float4 of the row
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = 16 * warp + 4 * it + lg;
      const int dst = sDst[r];
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), hv = gv;
      if (dst >= 0) {
        const float m = sMult[r];
        gv = __ldg(reinterpret_cast<const float4*>(g + (int64_t)dst * D) + q);
        hv = __ldg(reinterpret_cast<const float4*>(h + (int64_t)sSrc[r] * D) + q);
        gv.x *= m, gv.y *= m, gv.z *= m, gv.w *= m;
      }
      *reinterpret_cast<float4*>(&sG[r * DT_LD + 4 * q]) = gv;
      *reinterpret_cast<float4*>(&sH[r * DT_LD + 4 * q]) = hv;
    }
    __syncthreads();
    // group grp accumulates entries [32 grp, 32 grp + 32) of the pass (rows beyond n are zero)
#pragma unroll 4
    for (int r = 32 * grp; r < 32 * grp + 32; ++r) {
      const float4 gv = *reinterpret_cast<const float4*>(&sG[r * DT_LD + 4 * lq]);
      const float4 hv = *reinterpret_cast<const float4*>(&sH[r * DT_LD + 4 * mq]);
      const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, hb[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ga[a], hb[b], acc[a][b]);
    }
    __syncthreads();
  }
  // combine the four groups in group order through the (now idle) staging arrays: plain 32 x 32 blocks, two per array
  static_assert(2 * D * D <= DT_SUB * DT_LD, "reduction scratch");
  float* red2 = grp < 2 ? sG + grp * D * D : sH + (grp - 2) * D * D;
#pragma unroll
  for (int a = 0; a < 4; ++a)
    *reinterpret_cast<float4*>(&red2[(4 * lq + a) * D + 4 * mq]) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
  __syncthreads();
  {
    const float4 p0 = reinterpret_cast<const float4*>(sG)[tid], p1 = reinterpret_cast<const float4*>(sG + D * D)[tid];
    const float4 p2 = reinterpret_cast<const float4*>(sH)[tid], p3 = reinterpret_cast<const float4*>(sH + D * D)[tid];
    float4 o;
    o.x = ((p0.x + p1.x) + p2.x) + p3.x, o.y = ((p0.y + p1.y) + p2.y) + p3.y;
    o.z = ((p0.z + p1.z) + p2.z) + p3.z, o.w = ((p0.w + p1.w) + p2.w) + p3.w;
    reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * D * D)[tid] = o;
  }
}

// dTable[bucket][lm] = sum over the bucket's chunks (index order)
__global__ void dtable_reduce_kernel(const int* __restrict__ bucket_chunk_ptr, const float* __restrict__ partial, int dd,
                                     float* __restrict__ dtable) {
  const int b = blockIdx.x;
  const int c0 = bucket_chunk_ptr[b], c1 = bucket_chunk_ptr[b + 1];
  for (int j = threadIdx.x; j < dd; j += blockDim.x) {
    float s = 0.f;
    for (int c = c0; c < c1; ++c) s += partial[(int64_t)c * dd + j];
    dtable[(int64_t)b * dd + j] = s;
  }
}

// dW[k][lm] = sum_b c[b][k] dTable[b][lm]   (one launch per tower; grid = (ceil(dd/256), K))
__global__ void dbond_w_kernel(const float* __restrict__ bond_emb, const float* __restrict__ dtable, int V, int K, int dd,
                               float* __restrict__ dW) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (j >= dd) return;
  float s = 0.f;
  for (int b = 0; b < V; ++b) s = fmaf(bond_emb[(int64_t)b * K + k], dtable[(int64_t)b * dd + j], s);
  dW[(int64_t)k * dd + j] = s;
}

// dbond_emb[b][k] += <dTable_cat[b], Wc[k]> + <dTable_an[b], Wa[k]>   (grid = (V, ceil(K/8)); one warp per (b, k))
__global__ void dbond_emb_kernel(const float* __restrict__ dtable_cat, const float* __restrict__ dtable_an,
                                 const float* __restrict__ W_cat, const float* __restrict__ W_an, int V, int K, int dd,
                                 float* __restrict__ dbond_emb) {
  const int b = blockIdx.x, k = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= K) return;
  float s = 0.f;
  for (int j = lane; j < dd; j += 32) s = fmaf(dtable_cat[(int64_t)b * dd + j], W_cat[(int64_t)k * dd + j], s);
  for (int j = lane; j < dd; j += 32) s = fmaf(dtable_an[(int64_t)b * dd + j], W_an[(int64_t)k * dd + j], s);
  float tot = 0.f;
  for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, s, l);  // fixed order
  if (lane == 0) dbond_emb[(int64_t)b * K + k] += tot;
}

// ================================================================================================ B1
constexpr int EB_WARPS = 8;
// Each warp owns a contiguous atom range and a private [vocab][d] table in shared memory (no atomics); the CTA sums
// its warps' tables in warp order and writes partial[cta][vocab*d].
__global__ void __launch_bounds__(EB_WARPS * 32) embed_bwd_kernel(const int* __restrict__ atom_id, const float* __restrict__ dh0,
                                                                  int n_atoms, int vocab, int d, int atoms_per_warp,
                                                                  float* __restrict__ partial) {
  extern __shared__ __align__(16) float tab[];  // [EB_WARPS][vocab * d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = tab + (int64_t)warp * vocab * d;
  for (int i = lane; i < vocab * d; i += 32) mine[i] = 0.f;
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * EB_WARPS + warp;
  const int64_t v0 = gw * atoms_per_warp, v1 = min((int64_t)n_atoms, v0 + atoms_per_warp);
  if (d == 32) {  // one column per lane: eight atoms' rows in flight, accumulated in atom order (bit-reproducible)
    for (int64_t v = v0; v < v1; v += 8) {
      int id[8];
      float g[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = v + u < v1;
        id[u] = ok ? min(max(__ldg(atom_id + v + u), 0), vocab - 1) : 0;
        g[u] = ok ? __ldg(dh0 + (v + u) * 32 + lane) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) mine[id[u] * 32 + lane] += g[u];
    }
  } else {
    for (int64_t v = v0; v < v1; ++v) {
      const int id = min(max(__ldg(atom_id + v), 0), vocab - 1);
      for (int j = lane; j < d; j += 32) mine[id * d + j] += __ldg(dh0 + v * d + j);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < vocab * d; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < EB_WARPS; ++w) s += tab[(int64_t)w * vocab * d + i];
    partial[(int64_t)blockIdx.x * vocab * d + i] = s;
  }
}

// ================================================================================================ optimizer
// norms2[v] = sum of squares of (gs * grad + l2[v] * 2 * param) over variable v (one CTA per variable, fixed order tree);
// gs = 1 / *pair_count when the bucket holds sum-gradients (imp_clip_adam_sparse), else 1.
// regs (optional, [n_vars]): l2[v] * sum(param^2), the variable's term of the regularised loss.
__global__ void __launch_bounds__(256) var_sumsq_kernel(const float* __restrict__ grad, const float* __restrict__ param,
                                                        const int64_t* __restrict__ var_off, const float* __restrict__ var_l2,
                                                        float* __restrict__ norms2, const float* __restrict__ pair_count,
                                                        float* __restrict__ regs) {
  __shared__ float red[256], red2[256];
  const int v = blockIdx.x;
  const int64_t o0 = var_off[v], o1 = var_off[v + 1];
  const float l2 = 2.0f * var_l2[v];
  const float gs = pair_count ? 1.0f / pair_count[0] : 1.0f;
  float s = 0.f, w2 = 0.f;
  for (int64_t i = o0 + threadIdx.x; i < o1; i += 256) {
    const float w = param[i];
    const float gq = fmaf(l2, w, grad ? grad[i] * gs : 0.f);
    s = fmaf(gq, gq, s);
    w2 = fmaf(w, w, w2);
  }
  red[threadIdx.x] = s, red2[threadIdx.x] = w2;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w], red2[threadIdx.x] += red2[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (norms2) norms2[v] = red[0];
    if (regs) regs[v] = var_l2[v] * red2[0];
  }
}

// Variables listed in occ_var take the per-occurrence squared norm (csrc/occ_norm.cu) instead of the dense one.
__global__ void occ_norm_override_kernel(float* __restrict__ norms2, const int32_t* __restrict__ occ_var,
                                         const float* __restrict__ occ_norm2, int n_occ, const float* __restrict__ pair_count) {
  const int i = threadIdx.x;
  const float gs = pair_count ? 1.0f / pair_count[0] : 1.0f;
  if (i < n_occ) norms2[occ_var[i]] = occ_norm2[i] * gs * gs;
}

// loss = sse * (1 / count) + sum_v regs[v] (serial, fixed order: n_vars is ~100)
__global__ void loss_kernel(const float* __restrict__ sse, const float* __restrict__ pair_count, float inv_batch,
                            const float* __restrict__ regs, int n_vars, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float r = 0.f;
    for (int v = 0; v < n_vars; ++v) r += regs[v];
    loss[0] = sse[0] * (pair_count ? 1.0f / pair_count[0] : inv_batch) + r;
  }
}

// sum of squared errors of n predictions (one CTA, fixed order) -- evaluate()
__global__ void __launch_bounds__(256) sse_kernel(const float* __restrict__ pred, const float* __restrict__ y, int64_t n,
                                                  float* __restrict__ out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const float dlt = pred[i] - y[i];
    s = fmaf(dlt, dlt, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// [Keras semantics] per-variable clip_by_norm, then Adam: m += (g-m)(1-b1); v += (g^2-v)(1-b2);
// w -= alpha * m / (sqrt(v) + eps), alpha = lr sqrt(1-b2^t)/(1-b1^t) computed by the host.
__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                        float* __restrict__ m, float* __restrict__ vv,
                                                        const int64_t* __restrict__ var_off, const float* __restrict__ var_l2,
                                                        const float* __restrict__ norms2, int n_vars, float clipnorm, float alpha,
                                                        float beta1, float beta2, float eps, const float* __restrict__ pair_count) {
  const int v = blockIdx.x;
  const int64_t o0 = var_off[v], o1 = var_off[v + 1];
  const float l2 = 2.0f * var_l2[v];
  const float gs = pair_count ? 1.0f / pair_count[0] : 1.0f;
  const float nrm = sqrtf(norms2[v]);
  const float scale = clipnorm > 0.f ? clipnorm / fmaxf(nrm, clipnorm) : 1.0f;
  for (int64_t i = o0 + (int64_t)blockIdx.y * 256 + threadIdx.x; i < o1; i += (int64_t)gridDim.y * 256) {
    const float w = param[i];
    const float g = fmaf(l2, w, grad[i] * gs) * scale;
    const float mi = m[i] + (g - m[i]) * (1.0f - beta1);
    const float vi = vv[i] + (g * g - vv[i]) * (1.0f - beta2);
    m[i] = mi, vv[i] = vi;
    param[i] = w - alpha * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace imp

// =========================================================================================================== ABI
using namespace imp;

static int bw_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

extern "C" int imp_pool_bwd(const int32_t* d_mol_ptr, const int32_t* d_atom_id, int32_t n_mols, const float* d_dpooled, int32_t d,
                            float* d_dh, void* stream) {
  IMP_REQUIRE(n_mols >= 0 && d > 0, IMP_ERR_ARG, "imp_pool_bwd: bad sizes");
  if (n_mols == 0) return 0;
  IMP_REQUIRE(d_mol_ptr && d_atom_id && d_dpooled && d_dh, IMP_ERR_ARG, "imp_pool_bwd: null pointer");
  pool_bwd_kernel<<<(unsigned)ceil_div(n_mols, 8), 256, 0, (cudaStream_t)stream>>>(d_mol_ptr, d_atom_id, n_mols, d_dpooled, d, d_dh);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t imp_readout_bwd_workspace_floats(int32_t d, int32_t fp, int32_t mix, int32_t fp2) {
  return (int64_t)bw_sm_count() * RB_WARPS * (readout_grad_floats(d, fp, mix, fp2) + 1);
}

extern "C" int imp_readout_bwd(const float* d_pooled, int32_t n_pairs, int32_t d, int32_t fp, int32_t mix, int32_t fp2,
                               const imp_readout_weights_t* w_cat, const imp_readout_weights_t* w_an, const float* d_W1,
                               const float* d_b1, const float* d_W2, const float* d_b2, const float* d_T, const float* d_y,
                               float scale, float* d_dpooled, float* d_out, float* d_grads, float* d_sse, float* d_workspace,
                               void* stream) {
  IMP_REQUIRE(n_pairs >= 0, IMP_ERR_ARG, "imp_readout_bwd: negative size");
  IMP_REQUIRE(d >= 1 && fp >= 1 && mix >= 1 && d <= RB_MAXV && fp <= RB_MAXV && mix <= RB_MAXV && fp2 >= 0 && fp2 <= RB_MAXV,
              IMP_ERR_DIM, "imp_readout_bwd: d/fp/mix/fp2 = %d/%d/%d/%d must be in 1..%d", d, fp, mix, fp2, RB_MAXV);
  IMP_REQUIRE(d_pooled && d_y && d_dpooled && d_grads && d_sse && d_workspace && w_cat && w_an && d_W1 && d_b1 &&
                  (fp2 > 0 ? (d_W2 && d_b2) : d_T != nullptr),
              IMP_ERR_ARG, "imp_readout_bwd: null pointer");
  ReadoutBwdArgs a;
  a.pooled = d_pooled, a.T = d_T, a.y = d_y, a.wc = *w_cat, a.wa = *w_an, a.W1 = d_W1, a.b1 = d_b1, a.W2 = d_W2, a.b2 = d_b2;
  a.n_pairs = n_pairs, a.d = d, a.fp = fp, a.mix = mix, a.fp2 = fp2, a.scale = scale, a.d_pooled = d_dpooled, a.out = d_out;
  a.partial = d_workspace;
  const int nh = fp2 > 0 ? fp2 : 3;
  const int grid = bw_sm_count();
  const size_t smem = sizeof(float) * (2 * (2 * d * fp + fp + 2 * fp * mix + mix) + 2 * mix * nh + nh + (fp2 > 0 ? fp2 : 0) +
                                       RB_WARPS * 8 * RB_MAXV);
  IMP_CUDA(cudaFuncSetAttribute(readout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  IMP_REQUIRE(smem <= 96 * 1024, IMP_ERR_DIM, "imp_readout_bwd: needs %zu B of shared memory", smem);
  readout_bwd_kernel<<<grid, RB_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
  IMP_LAUNCH_CHECK();
  const int n = readout_grad_floats(d, fp, mix, fp2);
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_workspace, grid * RB_WARPS, n + 1, n, d_grads, 0);
  IMP_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_workspace + n, grid * RB_WARPS, n + 1, 1, d_sse, 0);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t imp_gated_update_bwd_workspace_floats(int32_t d) {
  return d == 32 ? (int64_t)bw_sm_count() * gru_grad_floats<32>() : (int64_t)IMP_ERR_DIM;
}

static int gated_update_bwd_any(const float* d_h, const float* d_agg, const float* d_z, const float* d_r, const float* d_ht,
                               const float* d_gout, int32_t n_atoms, int32_t n_cat_atoms, int32_t d, const imp_gru_weights_t* w_cat,
                               const imp_gru_weights_t* w_an, float eps, float* d_dh, float* d_dagg, float* d_grads_cat,
                               float* d_grads_an, float* d_workspace, void* stream, const char* who) {
  IMP_REQUIRE(n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG, "%s: bad sizes", who);
  IMP_REQUIRE(d == 32, IMP_ERR_DIM, "%s: atom_dim %d not supported (32)", who, d);
  IMP_REQUIRE(d_h && d_agg && d_gout && d_dh && d_dagg && d_grads_cat && d_grads_an && d_workspace && w_cat && w_an, IMP_ERR_ARG,
              "%s: null pointer", who);
  constexpr int D = 32;
  const int sms = bw_sm_count();
  const int tiles_cat = (int)ceil_div(n_cat_atoms, GB_TILE), tiles_an = (int)ceil_div(n_atoms - n_cat_atoms, GB_TILE);
  int n_cat = tiles_cat + tiles_an > 0 ? (int)((int64_t)sms * tiles_cat / (tiles_cat + tiles_an)) : 1;
  n_cat = n_cat < 1 ? 1 : (n_cat > sms - 1 ? sms - 1 : n_cat);
  const int grid = sms;
  const size_t smem = sizeof(GruBwdSmem<D>);
  if (d_z) {
    IMP_CUDA(cudaFuncSetAttribute(gated_update_bwd_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gated_update_bwd_kernel<D, true><<<grid, GB_THREADS, smem, (cudaStream_t)stream>>>(d_h, d_agg, d_gout, n_atoms, n_cat_atoms, n_cat,
                                                                                   *w_cat, *w_an, eps, d_dh, d_dagg, d_workspace,
                                                                                   d_z, d_r, d_ht);
  } else {
    IMP_CUDA(cudaFuncSetAttribute(gated_update_bwd_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gated_update_bwd_kernel<D, false><<<grid, GB_THREADS, smem, (cudaStream_t)stream>>>(d_h, d_agg, d_gout, n_atoms, n_cat_atoms, n_cat,
                                                                                    *w_cat, *w_an, eps, d_dh, d_dagg, d_workspace,
                                                                                    nullptr, nullptr, nullptr);
  }
  IMP_LAUNCH_CHECK();
  const int n = gru_grad_floats<D>();
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_workspace, n_cat, n, n, d_grads_cat, 0);
  IMP_LAUNCH_CHECK();
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_workspace + (int64_t)n_cat * n, grid - n_cat, n, n,
                                                                            d_grads_an, 0);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_gated_update_bwd(const float* d_h, const float* d_agg, const float* d_gout, int32_t n_atoms,
                                    int32_t n_cat_atoms, int32_t d, const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an,
                                    float eps, float* d_dh, float* d_dagg, float* d_grads_cat, float* d_grads_an,
                                    float* d_workspace, void* stream) {
  return gated_update_bwd_any(d_h, d_agg, nullptr, nullptr, nullptr, d_gout, n_atoms, n_cat_atoms, d, w_cat, w_an, eps, d_dh, d_dagg,
                              d_grads_cat, d_grads_an, d_workspace, stream, "imp_gated_update_bwd");
}

extern "C" int imp_gated_update_bwd_stored(const float* d_h, const float* d_agg, const float* d_z, const float* d_r, const float* d_ht,
                                           const float* d_gout, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                           const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_dh,
                                           float* d_dagg, float* d_grads_cat, float* d_grads_an, float* d_workspace, void* stream) {
  IMP_REQUIRE(d_z && d_r && d_ht, IMP_ERR_ARG, "imp_gated_update_bwd_stored: null gate pointers");
  return gated_update_bwd_any(d_h, d_agg, d_z, d_r, d_ht, d_gout, n_atoms, n_cat_atoms, d, w_cat, w_an, eps, d_dh, d_dagg, d_grads_cat,
                              d_grads_an, d_workspace, stream, "imp_gated_update_bwd_stored");
}

extern "C" int imp_bond_transform_bwd(const imp_graph_t* g, const int32_t* d_entry_dst, const int32_t* d_chunk_begin,
                                      const int32_t* d_chunk_end, int32_t n_chunks, const int32_t* d_bucket_chunk_ptr,
                                      const float* d_dagg, const float* d_h, int32_t d, int32_t bond_dim, const float* d_bond_emb,
                                      const float* d_W_cat, const float* d_W_an, float* d_dW_cat, float* d_dW_an,
                                      float* d_dbond_emb /* accumulated */, float* d_dtable /* [2 V_b, d, d] scratch */,
                                      float* d_workspace /* [n_chunks, d, d] */, void* stream) {
  IMP_REQUIRE(g && g->bond_vocab > 0 && n_chunks >= 0, IMP_ERR_ARG, "imp_bond_transform_bwd: bad arguments");
  IMP_REQUIRE(d == 32, IMP_ERR_DIM, "imp_bond_transform_bwd: atom_dim %d not supported (32)", d);
  IMP_REQUIRE(d_entry_dst && d_chunk_begin && d_chunk_end && d_bucket_chunk_ptr && d_dagg && d_h && d_bond_emb && d_W_cat &&
                  d_W_an && d_dW_cat && d_dW_an && d_dbond_emb && d_dtable && (n_chunks == 0 || d_workspace),
              IMP_ERR_ARG, "imp_bond_transform_bwd: null pointer");
  constexpr int D = 32;
  const int V = g->bond_vocab, dd = D * D;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_chunks > 0) {
    dtable_partial_kernel<D><<<n_chunks, DT_CHUNK_THREADS, 0, st>>>(d_chunk_begin, d_chunk_end, g->bucket_perm, d_entry_dst,
                                                                    g->col_src, g->edge_bm, d_dagg, d_h, d_workspace);
    IMP_LAUNCH_CHECK();
  }
  dtable_reduce_kernel<<<2 * V, 256, 0, st>>>(d_bucket_chunk_ptr, d_workspace, dd, d_dtable);
  IMP_LAUNCH_CHECK();
  const dim3 gw((dd + 255) / 256, bond_dim);
  dbond_w_kernel<<<gw, 256, 0, st>>>(d_bond_emb, d_dtable, V, bond_dim, dd, d_dW_cat);
  IMP_LAUNCH_CHECK();
  dbond_w_kernel<<<gw, 256, 0, st>>>(d_bond_emb, d_dtable + (int64_t)V * dd, V, bond_dim, dd, d_dW_an);
  IMP_LAUNCH_CHECK();
  const dim3 ge(V, (bond_dim + 7) / 8);
  dbond_emb_kernel<<<ge, 256, 0, st>>>(d_dtable, d_dtable + (int64_t)V * dd, d_W_cat, d_W_an, V, bond_dim, dd, d_dbond_emb);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t imp_embed_bwd_workspace_floats(int32_t atom_vocab, int32_t d) { return (int64_t)bw_sm_count() * atom_vocab * d; }

extern "C" int imp_embed_bwd(const int32_t* d_atom_id, const float* d_dh0, int32_t n_atoms, int32_t atom_vocab, int32_t d,
                             float* d_datom_emb, float* d_workspace, void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && atom_vocab > 0 && d > 0, IMP_ERR_ARG, "imp_embed_bwd: bad sizes");
  IMP_REQUIRE(d_atom_id && d_dh0 && d_datom_emb && d_workspace, IMP_ERR_ARG, "imp_embed_bwd: null pointer");
  const size_t smem = sizeof(float) * EB_WARPS * atom_vocab * d;
  IMP_REQUIRE(smem <= 200 * 1024, IMP_ERR_DIM, "imp_embed_bwd: vocabulary %d x dim %d does not fit shared memory", atom_vocab, d);
  const int grid = bw_sm_count();
  const int per_warp = (int)ceil_div(n_atoms, (int64_t)grid * EB_WARPS);
  IMP_CUDA(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  embed_bwd_kernel<<<grid, EB_WARPS * 32, smem, (cudaStream_t)stream>>>(d_atom_id, d_dh0, n_atoms, atom_vocab, d,
                                                                        per_warp > 0 ? per_warp : 1, d_workspace);
  IMP_LAUNCH_CHECK();
  const int n = atom_vocab * d;
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_workspace, grid, n, n, d_datom_emb, 0);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_clip_adam(float* d_param, const float* d_grad, float* d_m, float* d_v, const int64_t* d_var_off,
                             const float* d_var_l2, int32_t n_vars, float* d_norms2, float clipnorm, float lr, float beta1,
                             float beta2, float eps, int32_t step, void* stream) {
  IMP_REQUIRE(n_vars >= 0 && step >= 1, IMP_ERR_ARG, "imp_clip_adam: bad arguments");
  if (n_vars == 0) return 0;
  IMP_REQUIRE(d_param && d_grad && d_m && d_v && d_var_off && d_var_l2 && d_norms2, IMP_ERR_ARG, "imp_clip_adam: null pointer");
  var_sumsq_kernel<<<n_vars, 256, 0, (cudaStream_t)stream>>>(d_grad, d_param, d_var_off, d_var_l2, d_norms2, nullptr, nullptr);
  IMP_LAUNCH_CHECK();
  const float alpha = (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
  clip_adam_kernel<<<dim3(n_vars, 8), 256, 0, (cudaStream_t)stream>>>(d_param, d_grad, d_m, d_v, d_var_off, d_var_l2, d_norms2,
                                                                      n_vars, clipnorm, alpha, beta1, beta2, eps, nullptr);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_clip_adam_sparse(float* d_param, const float* d_grad, float* d_m, float* d_v, const int64_t* d_var_off,
                                    const float* d_var_l2, int32_t n_vars, float* d_norms2, float clipnorm, float lr, float beta1,
                                    float beta2, float eps, int32_t step, int32_t n_occ, const int32_t* d_occ_var,
                                    const float* d_occ_norm2, const float* d_pair_count, const float* d_sse, float inv_batch,
                                    float* d_loss, void* stream) {
  IMP_REQUIRE(n_vars >= 0 && step >= 1 && n_occ >= 0 && n_occ <= 32, IMP_ERR_ARG, "imp_clip_adam_sparse: bad arguments");
  if (n_vars == 0) return 0;
  IMP_REQUIRE(d_param && d_grad && d_m && d_v && d_var_off && d_var_l2 && d_norms2 && (n_occ == 0 || (d_occ_var && d_occ_norm2)),
              IMP_ERR_ARG, "imp_clip_adam_sparse: null pointer");
  IMP_REQUIRE(!d_loss || d_sse, IMP_ERR_ARG, "imp_clip_adam_sparse: d_loss needs d_sse");
  cudaStream_t st = (cudaStream_t)stream;
  // d_norms2 holds 2 * n_vars floats: squared norms, then the l2 loss terms of the variables
  var_sumsq_kernel<<<n_vars, 256, 0, st>>>(d_grad, d_param, d_var_off, d_var_l2, d_norms2, d_pair_count, d_norms2 + n_vars);
  IMP_LAUNCH_CHECK();
  if (n_occ > 0) {
    occ_norm_override_kernel<<<1, 32, 0, st>>>(d_norms2, d_occ_var, d_occ_norm2, n_occ, d_pair_count);
    IMP_LAUNCH_CHECK();
  }
  if (d_loss) {  // the loss of THIS step's forward pass (weights before the update), like Keras reports it
    loss_kernel<<<1, 32, 0, st>>>(d_sse, d_pair_count, inv_batch, d_norms2 + n_vars, n_vars, d_loss);
    IMP_LAUNCH_CHECK();
  }
  const float alpha = (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
  clip_adam_kernel<<<dim3(n_vars, 8), 256, 0, st>>>(d_param, d_grad, d_m, d_v, d_var_off, d_var_l2, d_norms2, n_vars, clipnorm, alpha,
                                                    beta1, beta2, eps, d_pair_count);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_eval_loss(const float* d_pred, const float* d_y, int64_t n, const float* d_param, const int64_t* d_var_off,
                             const float* d_var_l2, int32_t n_vars, float* d_scratch, float* d_loss, void* stream) {
  IMP_REQUIRE(n > 0 && n_vars >= 0 && d_pred && d_y && d_scratch && d_loss && (n_vars == 0 || (d_param && d_var_off && d_var_l2)),
              IMP_ERR_ARG, "imp_eval_loss: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: [0] sse, [1 .. 1 + n_vars) l2 terms
  sse_kernel<<<1, 256, 0, st>>>(d_pred, d_y, n, d_scratch);
  IMP_LAUNCH_CHECK();
  if (n_vars > 0) {
    var_sumsq_kernel<<<n_vars, 256, 0, st>>>(nullptr, d_param, d_var_off, d_var_l2, nullptr, nullptr, d_scratch + 1);
    IMP_LAUNCH_CHECK();
  }
  loss_kernel<<<1, 32, 0, st>>>(d_scratch, nullptr, 1.0f / (float)n, d_scratch + 1, n_vars, d_loss);
  IMP_LAUNCH_CHECK();
  return 0;
}
