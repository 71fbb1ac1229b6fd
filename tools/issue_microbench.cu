// Issue-rate microbenchmark for the fused forward's instruction mix (debug aid): warp-instructions per clock per SM of
// HFMA2, FFMA, MUFU.TANH, F2FP and a 4:1 HFMA2:LDS.128 mix, at 16 warps per SM (the fused kernel's occupancy) and 32.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_prof/issue_microbench tools/issue_microbench.cu && tools/_prof/issue_microbench
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void k(float* out, long long* cyc, int n) {
  __shared__ uint4 sm[512];
  sm[threadIdx.x % 512] = make_uint4(threadIdx.x, 1, 2, 3);
  __syncthreads();
  __half2 a[16];
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = __floats2half2_rn(0.001f * i, 0.002f * i), f[i] = 0.001f * i + threadIdx.x;
  __half2 x = __floats2half2_rn(1.0001f, 0.9999f), y = __floats2half2_rn(0.0001f, 0.0002f);
  float fx = 1.0001f, fy = 0.0001f;
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = i;
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __hfma2(a[i], x, y);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], fx, fy);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        f[2 * i] += __low2float(t);
      }
    } else if (MODE == 4) {  // 16 HFMA2 : 1 LDS.128 (the Z build's ratio)
      const uint4 v = sm[(threadIdx.x * 7 + it) & 511];
      x = *reinterpret_cast<const __half2*>(&v.x);
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __hfma2(a[i], x, y);
    } else if (MODE == 5) {  // 8 HFMA2 + 8 FFMA interleaved
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __hfma2(a[i], x, y), f[i] = fmaf(f[i], fx, fy);
    } else if (MODE == 6) {  // 8 FFMA + 8 integer ALU (LOP3/IADD)
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], fx, fy), u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u;
    } else if (MODE == 7) {  // 12 FFMA + 4 MUFU (gate-math ratio)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
        f[4 + 3 * i] = fmaf(f[4 + 3 * i], fx, fy), f[5 + 3 * i] = fmaf(f[5 + 3 * i], fx, fy), f[6 + 3 * i] = fmaf(f[6 + 3 * i], fx, fy);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i] + __low2float(a[i]) + __high2float(a[i]);
  for (int i = 0; i < 8; ++i) s += u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int threads) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, threads>>>(out, cyc, 64);
  k<MODE><<<148, threads>>>(out, cyc, ITERS);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double winst = (double)ITERS * per_iter * (threads / 32);
  printf("%-34s %4d threads/SM: %.3f warp-instr/clk/SM (%.3f per SMSP)\n", name, threads, winst / avg, winst / avg / 4);
  cudaFree(out), cudaFree(cyc);
}

int main() {
  for (int th : {512, 1024}) {
    run<0>("HFMA2", 16, th);
    run<1>("FFMA", 16, th);
    run<2>("MUFU.TANH", 16, th);
    run<3>("F2FP + unpack + FADD (3 instr)", 24, th);
    run<4>("16 HFMA2 : 1 LDS.128", 17, th);
    run<5>("HFMA2 + FFMA 1:1", 16, th);
    run<6>("FFMA + int ALU (3) 1:3", 32, th);
    run<7>("12 FFMA : 4 MUFU", 16, th);
  }
  return 0;
}
