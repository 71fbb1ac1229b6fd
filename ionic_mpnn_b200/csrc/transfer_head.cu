// Layers of the transfer-learning head of train_melting_point_transfer.py:95-104 and their backward passes:
//   Dense(256, relu) -> BatchNormalization -> Dense(128, relu) -> Dropout(0.3) -> Dense(64, relu) -> Dense(1),
//   loss = tf.keras.losses.Huber(delta=1.0) (:196, :224).
// The head sees one row per ion PAIR (the output of "mix_cat_an"), a few thousand rows at most per training batch: every
// reduction here is one thread walking the rows in order (bit-reproducible), not a tuned GEMM.
// Keras semantics restated where they matter:
//   * BatchNormalization on a 2-D input takes the non-fused path: batch moments are the BIASED variance (tf.nn.moments), the
//     moving averages are updated as moving * momentum + batch * (1 - momentum) with that same variance; defaults
//     momentum 0.99, epsilon 1e-3.
//   * Dropout scales the kept elements by 1 / (1 - rate) at training time and is the identity at inference.
//   * Huber: 0.5 e^2 for |e| <= delta, delta (|e| - 0.5 delta) otherwise; mean over the batch.
#include "common.cuh"

namespace imp {

// ---------------------------------------------------------------------------------------------- Dense backward
// gx[r][i] = sum_j g[r][j] W[i][j],  g = gy (* [y > 0] for relu)
__global__ void __launch_bounds__(256) dense_bwd_x_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ W,
                                                          int64_t rows, int in, int out, int relu, float* __restrict__ gx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * in) return;
  const int64_t r = idx / in;
  const int i = (int)(idx % in);
  float acc = 0.f;
  for (int j = 0; j < out; ++j) {
    float g = gy[r * out + j];
    if (relu && !(y[r * out + j] > 0.f)) g = 0.f;
    acc = fmaf(g, W[i * out + j], acc);
  }
  gx[idx] = acc;
}
// gW[i][j] = sum_r x[r][i] g[r][j];  gb[j] = sum_r g[r][j]   (thread (i, j); i == in computes the bias)
__global__ void __launch_bounds__(256) dense_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ y,
                                                          int64_t rows, int in, int out, int relu, float* __restrict__ gW,
                                                          float* __restrict__ gb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (in + 1) * out) return;
  const int i = idx / out, j = idx % out;
  float acc = 0.f;
  for (int64_t r = 0; r < rows; ++r) {
    float g = gy[r * out + j];
    if (relu && !(y[r * out + j] > 0.f)) g = 0.f;
    acc = i < in ? fmaf(x[r * in + i], g, acc) : acc + g;
  }
  if (i < in) gW[i * out + j] = acc;
  else if (gb) gb[j] = acc;
}

// ---------------------------------------------------------------------------------------------- BatchNormalization
__global__ void batchnorm_fwd_kernel(const float* __restrict__ x, int64_t rows, int c, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ mov_mean, float* __restrict__ mov_var,
                                     float momentum, float eps, int training, float* __restrict__ y, float* __restrict__ save_mean,
                                     float* __restrict__ save_inv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  float mean, var;
  if (training) {
    float s = 0.f;
    for (int64_t r = 0; r < rows; ++r) s += x[r * c + j];
    mean = s / (float)rows;
    float q = 0.f;
    for (int64_t r = 0; r < rows; ++r) {
      const float dlt = x[r * c + j] - mean;
      q = fmaf(dlt, dlt, q);
    }
    var = q / (float)rows;
    mov_mean[j] = mov_mean[j] * momentum + mean * (1.f - momentum);
    mov_var[j] = mov_var[j] * momentum + var * (1.f - momentum);
  } else {
    mean = mov_mean[j], var = mov_var[j];
  }
  const float inv = 1.0f / sqrtf(var + eps);
  if (save_mean) save_mean[j] = mean, save_inv[j] = inv;
  const float gsc = gamma[j] * inv, sh = beta[j] - mean * gsc;
  for (int64_t r = 0; r < rows; ++r) y[r * c + j] = fmaf(x[r * c + j], gsc, sh);
}
// training-mode backward: gx = gamma inv / n (n gy - sum gy - xhat sum(gy xhat))
__global__ void batchnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, int64_t rows, int c,
                                     const float* __restrict__ gamma, const float* __restrict__ save_mean, const float* __restrict__ save_inv,
                                     float* __restrict__ gx, float* __restrict__ ggamma, float* __restrict__ gbeta) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  const float mean = save_mean[j], inv = save_inv[j];
  float sb = 0.f, sg = 0.f;
  for (int64_t r = 0; r < rows; ++r) {
    const float g = gy[r * c + j];
    sb += g;
    sg = fmaf(g, (x[r * c + j] - mean) * inv, sg);
  }
  ggamma[j] = sg, gbeta[j] = sb;
  const float k = gamma[j] * inv / (float)rows;
  for (int64_t r = 0; r < rows; ++r) {
    const float xh = (x[r * c + j] - mean) * inv;
    gx[r * c + j] = k * ((float)rows * gy[r * c + j] - sb - xh * sg);
  }
}

// ---------------------------------------------------------------------------------------------- Dropout
// keep(i) is a pure function of (seed, i): the backward pass and the test oracle recompute the same mask.
__device__ __forceinline__ float dropout_uniform(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
__global__ void dropout_kernel(const float* __restrict__ x, int64_t n, float rate, unsigned long long seed, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float scale = 1.0f / (1.0f - rate);
  y[i] = dropout_uniform(seed, (unsigned long long)i) >= rate ? x[i] * scale : 0.f;
}

// ---------------------------------------------------------------------------------------------- Huber
__global__ void __launch_bounds__(1024) huber_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n, float delta,
                                                     float scale, float* __restrict__ loss_sum, float* __restrict__ dpred) {
  __shared__ float red[1024];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float e = pred[i] - target[i], a = fabsf(e);
    acc += a <= delta ? 0.5f * e * e : delta * (a - 0.5f * delta);
    if (dpred) dpred[i] = scale * fminf(fmaxf(e, -delta), delta);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {  // fixed tree: bit-reproducible
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss_sum) *loss_sum = red[0];
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

}  // namespace imp

using namespace imp;

extern "C" int imp_dense_bwd(const float* d_x, const float* d_y, const float* d_gy, int64_t rows, int32_t in_dim, int32_t out_dim,
                             const float* d_kernel, int32_t activation, float* d_gx, float* d_gkernel, float* d_gbias, void* stream) {
  IMP_REQUIRE(rows >= 0 && in_dim >= 1 && out_dim >= 1, IMP_ERR_ARG, "imp_dense_bwd: bad sizes");
  IMP_REQUIRE(activation == 0 || activation == 1, IMP_ERR_ARG, "imp_dense_bwd: activation must be 0 (linear) or 1 (relu)");
  IMP_REQUIRE(d_gy && d_kernel && (activation == 0 || d_y), IMP_ERR_ARG, "imp_dense_bwd: null pointer (relu needs the layer's output)");
  IMP_REQUIRE(!d_gkernel || d_x, IMP_ERR_ARG, "imp_dense_bwd: the kernel gradient needs the layer's input");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_gx && rows > 0) {
    const int64_t n = rows * in_dim;
    dense_bwd_x_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_gy, d_y, d_kernel, rows, in_dim, out_dim, activation, d_gx);
    IMP_LAUNCH_CHECK();
  }
  if (d_gkernel) {
    dense_bwd_w_kernel<<<(unsigned)ceil_div((int64_t)(in_dim + 1) * out_dim, 256), 256, 0, st>>>(d_x, d_gy, d_y, rows, in_dim, out_dim,
                                                                                               activation, d_gkernel, d_gbias);
    IMP_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int imp_batchnorm(const float* d_x, int64_t rows, int32_t channels, const float* d_gamma, const float* d_beta,
                             float* d_moving_mean, float* d_moving_var, float momentum, float eps, int32_t training, float* d_y,
                             float* d_save_mean, float* d_save_inv, void* stream) {
  IMP_REQUIRE(rows >= 0 && channels >= 1, IMP_ERR_ARG, "imp_batchnorm: bad sizes");
  IMP_REQUIRE(d_x && d_gamma && d_beta && d_moving_mean && d_moving_var && d_y, IMP_ERR_ARG, "imp_batchnorm: null pointer");
  IMP_REQUIRE(!training || rows > 0, IMP_ERR_ARG, "imp_batchnorm: batch statistics of an empty batch");
  IMP_REQUIRE((d_save_mean == nullptr) == (d_save_inv == nullptr), IMP_ERR_ARG, "imp_batchnorm: pass both saved statistics or neither");
  batchnorm_fwd_kernel<<<(unsigned)ceil_div(channels, 128), 128, 0, (cudaStream_t)stream>>>(
      d_x, rows, channels, d_gamma, d_beta, d_moving_mean, d_moving_var, momentum, eps, training, d_y, d_save_mean, d_save_inv);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_batchnorm_bwd(const float* d_x, const float* d_gy, int64_t rows, int32_t channels, const float* d_gamma,
                                 const float* d_save_mean, const float* d_save_inv, float* d_gx, float* d_ggamma, float* d_gbeta,
                                 void* stream) {
  IMP_REQUIRE(rows >= 1 && channels >= 1, IMP_ERR_ARG, "imp_batchnorm_bwd: bad sizes");
  IMP_REQUIRE(d_x && d_gy && d_gamma && d_save_mean && d_save_inv && d_gx && d_ggamma && d_gbeta, IMP_ERR_ARG, "imp_batchnorm_bwd: null pointer");
  batchnorm_bwd_kernel<<<(unsigned)ceil_div(channels, 128), 128, 0, (cudaStream_t)stream>>>(d_x, d_gy, rows, channels, d_gamma, d_save_mean,
                                                                                          d_save_inv, d_gx, d_ggamma, d_gbeta);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_dropout(const float* d_x, int64_t n, float rate, uint64_t seed, float* d_y, void* stream) {
  IMP_REQUIRE(n >= 0 && rate >= 0.f && rate < 1.f, IMP_ERR_ARG, "imp_dropout: rate must be in [0, 1)");
  if (n == 0) return 0;
  IMP_REQUIRE(d_x && d_y, IMP_ERR_ARG, "imp_dropout: null pointer");
  dropout_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_x, n, rate, (unsigned long long)seed, d_y);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_huber(const float* d_pred, const float* d_target, int64_t n, float delta, float scale, float* d_loss_sum,
                         float* d_dpred, void* stream) {
  IMP_REQUIRE(n >= 0 && delta > 0.f, IMP_ERR_ARG, "imp_huber: bad arguments");
  IMP_REQUIRE(d_pred && d_target && (d_loss_sum || d_dpred), IMP_ERR_ARG, "imp_huber: null pointer");
  huber_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_pred, d_target, n, delta, scale, d_loss_sum, d_dpred);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_add(const float* d_a, const float* d_b, int64_t n, float* d_y, void* stream) {
  IMP_REQUIRE(n >= 0, IMP_ERR_ARG, "imp_add: bad size");
  if (n == 0) return 0;
  IMP_REQUIRE(d_a && d_b && d_y, IMP_ERR_ARG, "imp_add: null pointer");
  add_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, d_y);
  IMP_LAUNCH_CHECK();
  return 0;
}
