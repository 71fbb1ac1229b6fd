// keras.layers.Dense on [rows, in] device tensors: y = act(x . kernel + bias), kernel (in, out) row-major as Keras stores it
// (train_viscosity.py:189 Dense(fp_size, relu), :197-198 Dense(mixing_size, relu), :204 Dense(3); train_melting_point.py:173-198).
// The model path never calls this -- imp_pool_head_* / imp_readout_* fuse these layers with the pooling and the head; it exists
// so that a reference-style `encode()` written against the layer classes (ionic_mpnn_b200/layers.py) runs end to end.
// Rows are molecule-sized (P or 2P), in / out <= 1024: an HBM stream with the kernel matrix resident in shared memory.
#include "common.cuh"

namespace imp {

constexpr int DN_ROWS = 32;  // rows per CTA pass

__global__ void __launch_bounds__(256) dense_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                                    int64_t rows, int in, int out, int relu, float* __restrict__ y) {
  extern __shared__ float sm[];
  float* sW = sm;             // in * out
  float* sx = sm + in * out;  // DN_ROWS * in
  for (int i = threadIdx.x; i < in * out; i += 256) sW[i] = W[i];
  for (int64_t r0 = (int64_t)blockIdx.x * DN_ROWS; r0 < rows; r0 += (int64_t)gridDim.x * DN_ROWS) {
    const int nr = (int)min((int64_t)DN_ROWS, rows - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * in; i += 256) sx[i] = x[r0 * in + i];
    __syncthreads();
    for (int o = threadIdx.x; o < nr * out; o += 256) {
      const int r = o / out, j = o % out;
      float acc = b ? b[j] : 0.f;
      for (int i = 0; i < in; ++i) acc = fmaf(sx[r * in + i], sW[i * out + j], acc);  // fixed order: deterministic
      y[(r0 + r) * out + j] = relu ? fmaxf(acc, 0.f) : acc;
    }
  }
}

}  // namespace imp

using namespace imp;

extern "C" int imp_dense(const float* d_x, int64_t rows, int32_t in_dim, int32_t out_dim, const float* d_kernel, const float* d_bias,
                         int32_t activation, float* d_y, void* stream) {
  IMP_REQUIRE(rows >= 0 && in_dim >= 1 && out_dim >= 1, IMP_ERR_ARG, "imp_dense: bad sizes");
  IMP_REQUIRE(activation == 0 || activation == 1, IMP_ERR_ARG, "imp_dense: activation must be 0 (linear) or 1 (relu)");
  if (rows == 0) return 0;
  IMP_REQUIRE(d_x && d_kernel && d_y, IMP_ERR_ARG, "imp_dense: null pointer");
  const size_t smem = ((size_t)in_dim * out_dim + (size_t)DN_ROWS * in_dim) * sizeof(float);
  IMP_REQUIRE(smem <= 200 * 1024, IMP_ERR_DIM, "imp_dense: kernel matrix %d x %d does not fit shared memory", in_dim, out_dim);
  IMP_CUDA(cudaFuncSetAttribute(dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t want = ceil_div(rows, DN_ROWS);
  const int grid = (int)(want < 4 * 148 ? want : 4 * 148);
  dense_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(d_x, d_kernel, d_bias, rows, in_dim, out_dim, activation, d_y);
  IMP_LAUNCH_CHECK();
  return 0;
}
