// Microbenchmark (debug aid): issue rate of small tcgen05.mma (M = 128, K = 16, f16) as a function of N, of where the
// A operand lives (tensor memory vs shared memory) and of whether consecutive MMAs share an accumulator.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ionic_mpnn_b200/csrc -I include tools/mma_microbench.cu -o tools/_prof/mma_microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace imp;

__global__ void __launch_bounds__(128) bench(int N, int a_in_tmem, int n_acc, int n_mma, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h pairs
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  {
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = 0x3c003c00u;
    for (int c = 0; c < 128; c += 32) tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tc::tmem_wait_st();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    tc::fence_after_thread_sync();
    const uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, N);
    const uint32_t sA = tc::smem_u32(smem), sB = tc::smem_u32(smem) + 16384;
    t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + 128 + (i % n_acc) * N;
      const uint64_t bd = tc::make_smem_desc(sB + (i % 8) * 2 * N * 16, N * 16, 128);
      if (a_in_tmem)
        tc::mma_f16_ts(d, tmem + 8 * (i % 16), bd, idesc, i >= n_acc);
      else
        tc::mma_bf16(d, tc::make_smem_desc(sA + (i % 4) * 4096, 2048, 128), bd, idesc, i >= n_acc);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  if (tid == 0) { t1 = clock64(); out[0] = t1 - t0; }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("%4s %8s %6s %6s %10s %10s\n", "N", "A", "n_acc", "n_mma", "cycles", "cyc/mma");
  for (int N : {32, 64, 96, 128, 256})
    for (int a_in_tmem : {1, 0})
      for (int n_acc : {1, 4})
        for (int n_mma : {16, 64}) {
          if (n_acc * N > 384) continue;
          long long best = 1ll << 60;
          for (int rep = 0; rep < 5; ++rep) {
            bench<<<1, 128, 64 * 1024>>>(N, a_in_tmem, n_acc, n_mma, d);
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            if (h < best) best = h;
          }
          printf("%4d %8s %6d %6d %10lld %10.1f\n", N, a_in_tmem ? "tmem" : "smem", n_acc, n_mma, best, (double)best / n_mma);
        }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
