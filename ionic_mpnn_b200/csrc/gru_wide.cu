// GatedUpdate.call (models/layers.py:142-156) for wide atom states (atom_dim 128 / 256: BASELINE configs[4], the
// "wide/deep" variant), fp32.  At these widths the three Dense(2d -> d) layers no longer fit the one-thread-per-atom
// scheme of gated_update_kernel (weights 1.5 MB at d = 256), so the layer runs as three tiled SIMT GEMMs with fused
// gate epilogues plus one LayerNorm/residual kernel.  This is the general-shape fp32 path (parity first); the
// tensor-core version of this shape is future work.
#include <math.h>

#include "common.cuh"

namespace imp {

constexpr int GW_BM = 64, GW_BN = 64, GW_BK = 16;

// out[a][j] = act(bias[j] + sum_k x0[a][k] W[k][j] + sum_k x1[a][k] W[d + k][j]);  ACT 0: sigmoid, 1: sigmoid * gate_in, 2: tanh
template <int ACT>
__global__ void __launch_bounds__(256) gw_dense2_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                        const float* __restrict__ W, const float* __restrict__ bias,
                                                        const float* __restrict__ gate_in, int n_rows, int d,
                                                        float* __restrict__ out) {
  __shared__ float sA[GW_BK][GW_BM + 4];
  __shared__ float sB[GW_BK][GW_BN + 4];
  const int row0 = blockIdx.x * GW_BM, col0 = blockIdx.y * GW_BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < 2 * d; k0 += GW_BK) {
    const float* x = k0 < d ? x0 : x1;
    const int kx = k0 < d ? k0 : k0 - d;
    for (int i = threadIdx.x; i < GW_BM * GW_BK; i += 256) {
      const int r = i / GW_BK, c = i % GW_BK;
      sA[c][r] = row0 + r < n_rows ? x[(int64_t)(row0 + r) * d + kx + c] : 0.f;
    }
    for (int i = threadIdx.x; i < GW_BK * GW_BN; i += 256) {
      const int r = i / GW_BN, c = i % GW_BN;
      sB[r][c] = col0 + c < d ? W[(int64_t)(k0 + r) * d + col0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GW_BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c >= d) continue;
      const float pre = acc[i][j] + bias[c];
      float v;
      if (ACT == 2)
        v = tanhf(pre);
      else {
        v = 1.0f / (1.0f + expf(-pre));
        if (ACT == 1) v *= gate_in[(int64_t)r * d + c];
      }
      out[(int64_t)r * d + c] = v;
    }
  }
}

// one warp per atom: n = (1 - z) h + z ht; LayerNorm(eps, biased variance) * gamma + beta + h
__global__ void gw_finish_kernel(const float* __restrict__ h, const float* __restrict__ z, const float* __restrict__ ht,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, int n_rows, int d, float eps,
                                 float* __restrict__ out) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int64_t o = (int64_t)r * d;
  float sum = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float hj = h[o + j];
    sum += (1.0f - z[o + j]) * hj + z[o + j] * ht[o + j];
  }
  for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  const float mean = sum / d;
  float var = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float hj = h[o + j];
    const float c = (1.0f - z[o + j]) * hj + z[o + j] * ht[o + j] - mean;
    var = fmaf(c, c, var);
  }
  for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
  const float inv = 1.0f / sqrtf(var / d + eps);
  for (int j = lane; j < d; j += 32) {
    const float hj = h[o + j];
    const float n = (1.0f - z[o + j]) * hj + z[o + j] * ht[o + j];
    out[o + j] = (n - mean) * inv * gamma[j] + beta[j] + hj;
  }
}

static int gw_tower(const float* h, const float* agg, int n, int d, const imp_gru_weights_t& w, float eps, float* out, float* zb,
                    float* rhb, float* htb, cudaStream_t st) {
  if (n == 0) return 0;
  const dim3 grid((unsigned)ceil_div(n, GW_BM), (unsigned)ceil_div(d, GW_BN));
  gw_dense2_kernel<0><<<grid, 256, 0, st>>>(h, agg, w.Wz, w.bz, nullptr, n, d, zb);
  IMP_LAUNCH_CHECK();
  gw_dense2_kernel<1><<<grid, 256, 0, st>>>(h, agg, w.Wr, w.br, h, n, d, rhb);
  IMP_LAUNCH_CHECK();
  gw_dense2_kernel<2><<<grid, 256, 0, st>>>(rhb, agg, w.Wh, w.bh, nullptr, n, d, htb);
  IMP_LAUNCH_CHECK();
  gw_finish_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(h, zb, htb, w.gamma, w.beta, n, d, eps, out);
  IMP_LAUNCH_CHECK();
  return 0;
}

}  // namespace imp

using namespace imp;

extern "C" int64_t imp_gated_update_wide_workspace_floats(int32_t n_atoms, int32_t d) { return 3 * (int64_t)n_atoms * d; }

extern "C" int imp_gated_update_wide(const float* d_h, const float* d_agg, int32_t n_atoms, int32_t n_cat_atoms, int32_t d,
                                     const imp_gru_weights_t* w_cat, const imp_gru_weights_t* w_an, float eps, float* d_h_out,
                                     float* d_workspace, void* stream) {
  IMP_REQUIRE(n_atoms >= 0 && n_cat_atoms >= 0 && n_cat_atoms <= n_atoms, IMP_ERR_ARG, "imp_gated_update_wide: bad sizes");
  IMP_REQUIRE(d > 0 && d % GW_BK == 0, IMP_ERR_DIM, "imp_gated_update_wide: atom_dim %d must be a multiple of %d", d, GW_BK);
  if (n_atoms == 0) return 0;
  IMP_REQUIRE(d_h && d_agg && d_h_out && d_workspace && w_cat && w_an && w_cat->Wz && w_an->Wz, IMP_ERR_ARG,
              "imp_gated_update_wide: null pointer");
  const int64_t nd = (int64_t)n_atoms * d, off = (int64_t)n_cat_atoms * d;
  float *zb = d_workspace, *rhb = d_workspace + nd, *htb = d_workspace + 2 * nd;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = gw_tower(d_h, d_agg, n_cat_atoms, d, *w_cat, eps, d_h_out, zb, rhb, htb, st)) return rc;
  return gw_tower(d_h + off, d_agg + off, n_atoms - n_cat_atoms, d, *w_an, eps, d_h_out + off, zb + off, rhb + off, htb + off, st);
}
