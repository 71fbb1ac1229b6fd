// Diagnostic: one CTA computes D[128,N] = A[128,K] * B[N,K]^T with tcgen05.mma (bf16 or tf32 operands, fp32
// accumulate in TMEM) using exactly the staging layout / descriptors of tc_common.cuh.  tests/ compares it with a
// CPU product, so a descriptor or layout mistake is caught in isolation, before it can hide inside a fused kernel.
#include "common.cuh"
#include "tc_common.cuh"

namespace imp {

template <int KIND>  // 0 = bf16, 1 = tf32
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ D, int N, int K, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int EPC = KIND == 0 ? 8 : 4;  // elements per 16-byte chunk
  const int C = K / EPC;                  // chunks along K
  unsigned char* sA = smem;
  unsigned char* sB = smem + 128 * C * 16;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < 128 * C; i += 128) {
    const int r = i / C, c = i % C;
    const float* src = A + (size_t)r * K + c * EPC;
    uint4 v;
    if (KIND == 0) {
      v.x = tc::pack_bf16x2(src[0], src[1]), v.y = tc::pack_bf16x2(src[2], src[3]);
      v.z = tc::pack_bf16x2(src[4], src[5]), v.w = tc::pack_bf16x2(src[6], src[7]);
    } else {
      v = *reinterpret_cast<const uint4*>(src);
    }
    *reinterpret_cast<uint4*>(sA + tc::chunk_off(r, c, 128)) = v;
  }
  for (int i = tid; i < N * C; i += 128) {
    const int r = i / C, c = i % C;
    const float* src = B + (size_t)r * K + c * EPC;
    uint4 v;
    if (KIND == 0) {
      v.x = tc::pack_bf16x2(src[0], src[1]), v.y = tc::pack_bf16x2(src[2], src[3]);
      v.z = tc::pack_bf16x2(src[4], src[5]), v.w = tc::pack_bf16x2(src[6], src[7]);
    } else {
      v = *reinterpret_cast<const uint4*>(src);
    }
    *reinterpret_cast<uint4*>(sB + tc::chunk_off(r, c, N)) = v;
  }
  if (warp == 0) tc::tmem_alloc<64>(&tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(KIND == 0 ? tc::FMT_BF16 : tc::FMT_TF32, 128, N);
    const uint32_t lboA = 128 * 16, lboB = N * 16, sbo = 128;
    for (int s = 0; s < C / 2; ++s) {
      const uint32_t a_addr = tc::smem_u32(sA) + 2 * s * lboA, b_addr = tc::smem_u32(sB) + 2 * s * lboB;
      const uint64_t ad = swap_lbo_sbo ? tc::make_smem_desc(a_addr, sbo, lboA) : tc::make_smem_desc(a_addr, lboA, sbo);
      const uint64_t bd = swap_lbo_sbo ? tc::make_smem_desc(b_addr, sbo, lboB) : tc::make_smem_desc(b_addr, lboB, sbo);
      if (KIND == 0)
        tc::mma_bf16(tmem, ad, bd, idesc, s > 0);
      else
        tc::mma_tf32(tmem, ad, bd, idesc, s > 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_thread_sync();
  for (int n0 = 0; n0 < N; n0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + n0 + j] = v[j];
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

// "TS" form: the A operand comes from tensor memory.  Thread r converts row r of A to 16-bit pairs and writes them
// with tcgen05.st (column j of the A region = elements 2j, 2j+1); B is staged in shared memory as above.
template <int FMT>
__global__ void __launch_bounds__(128) tc_selftest_ts_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int C = K / 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * C; i += 128) {
    const int r = i / C, c = i % C;
    const float* src = B + (size_t)r * K + c * 8;
    uint4 v;
    v.x = tc::pack2<FMT>(src[0], src[1]), v.y = tc::pack2<FMT>(src[2], src[3]);
    v.z = tc::pack2<FMT>(src[4], src[5]), v.w = tc::pack2<FMT>(src[6], src[7]);
    *reinterpret_cast<uint4*>(smem + tc::chunk_off(r, c, N)) = v;
  }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t tA = tmem, tD = tmem + 128;
  for (int c0 = 0; c0 < K / 2; c0 += 16) {  // K/2 columns, 16 at a time
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = tc::pack2<FMT>(A[(size_t)tid * K + 2 * (c0 + i)], A[(size_t)tid * K + 2 * (c0 + i) + 1]);
    tc::tmem_st16(tA + lane_off + c0, r);
  }
  tc::tmem_wait_st();
  tc::fence_before_thread_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_thread_sync();
    const uint32_t idesc = tc::make_idesc(FMT, 128, N);
    for (int s = 0; s < K / 16; ++s)
      tc::mma_f16_ts(tD, tA + 8 * s, tc::make_smem_desc(tc::smem_u32(smem) + 2 * s * N * 16, N * 16, 128), idesc, s > 0);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_thread_sync();
  for (int n0 = 0; n0 < N; n0 += 32) {
    float v[32];
    tc::tmem_ld32(tD + lane_off + n0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + n0 + j] = v[j];
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// MN-major operands (kind 4, tf32): D[128, N] = At^T . Bt with At [K][128] and Bt [K][N] row-major in global memory, i.e. the
// M / N index is the contiguous one -- the form in which a thread that owns one K index (one atom) can stage its row with
// 16-byte stores.  Canonical no-swizzle MN-major layout (cute mma_traits_sm100.hpp, "UmmaDescriptor Major-MN", INTERLEAVE:
// ((T,1,m),(8,k)) : ((1,T,SBO),(1T,LBO)), T = 4 tf32 per 16 bytes):
//     byte offset(mn, k) = (mn % 4) * 4 + (k % 8) * 16 + (mn / 4) * SBO + (k / 8) * LBO
// core matrix = 8 k x 16 bytes; here SBO = 128 (core matrices of consecutive mn chunks follow each other) and
// LBO = (R / 4) * 128 (one group of 8 k after the other).  One kind::tf32 MMA consumes exactly one group of 8 k.
__global__ void __launch_bounds__(128) tc_selftest_mn_kernel(const float* __restrict__ At, const float* __restrict__ Bt,
                                                             float* __restrict__ D, int N, int K, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lboA = (128 / 4) * 128, lboB = (uint32_t)(N / 4) * 128, sbo = 128;
  unsigned char* sA = smem;
  unsigned char* sB = smem + (size_t)(K / 8) * lboA;
  const bool a_kmajor = swap_lbo_sbo & 2, b_kmajor = swap_lbo_sbo & 4;  // debug: stage that operand K-major instead
  for (int i = tid; i < K * 32; i += 128) {  // (k, chunk of 4 m)
    const int k = i / 32, c = i % 32;
    const uint4 v = *reinterpret_cast<const uint4*>(At + (size_t)k * 128 + 4 * c);
    if (!a_kmajor) {
      *reinterpret_cast<uint4*>(sA + (k % 8) * 16 + c * sbo + (k / 8) * lboA) = v;
    } else {  // element (m, k) -> chunk_off(m, k / 4, 128) + (k % 4) * 4
      const float vv[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float*>(sA + tc::chunk_off(4 * c + q, k / 4, 128) + (k % 4) * 4) = vv[q];
    }
  }
  for (int i = tid; i < K * (N / 4); i += 128) {
    const int k = i / (N / 4), c = i % (N / 4);
    const uint4 v = *reinterpret_cast<const uint4*>(Bt + (size_t)k * N + 4 * c);
    if (!b_kmajor) {
      *reinterpret_cast<uint4*>(sB + (k % 8) * 16 + c * sbo + (k / 8) * lboB) = v;
    } else {
      const float vv[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float*>(sB + tc::chunk_off(4 * c + q, k / 4, N) + (k % 4) * 4) = vv[q];
    }
  }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, N) | (a_kmajor ? 0u : (1u << 15)) | (b_kmajor ? 0u : (1u << 16));
    const bool swap = swap_lbo_sbo & 1;
    for (int s = 0; s < K / 8; ++s) {
      // K-major staging: K-step s = chunks 2 s, 2 s + 1 (LBO = R * 16 between chunks); MN-major: one group of 8 k per step
      const uint32_t a_addr = tc::smem_u32(sA) + (a_kmajor ? 2 * s * 128 * 16 : s * lboA);
      const uint32_t b_addr = tc::smem_u32(sB) + (b_kmajor ? 2 * s * N * 16 : s * lboB);
      const uint64_t ad = a_kmajor ? tc::make_smem_desc(a_addr, 128 * 16, 128)
                                   : (swap ? tc::make_smem_desc(a_addr, sbo, lboA) : tc::make_smem_desc(a_addr, lboA, sbo));
      const uint64_t bd = b_kmajor ? tc::make_smem_desc(b_addr, N * 16, 128)
                                   : (swap ? tc::make_smem_desc(b_addr, sbo, lboB) : tc::make_smem_desc(b_addr, lboB, sbo));
      tc::mma_tf32(tmem, ad, bd, idesc, s > 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_thread_sync();
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + n0 + j] = v[j];
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

}  // namespace imp

extern "C" int imp_tc_selftest(const float* d_A, const float* d_B, float* d_D, int32_t N, int32_t K, int32_t kind,
                               int32_t swap_lbo_sbo, void* stream) {
  using namespace imp;
  IMP_REQUIRE(d_A && d_B && d_D, IMP_ERR_ARG, "imp_tc_selftest: null pointer");
  if (kind == 4) {  // tf32, both operands MN-major: d_A = At [K][128], d_B = Bt [K][N]
    IMP_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K > 0 && K % 8 == 0, IMP_ERR_ARG, "imp_tc_selftest: MN-major form needs N %% 16 == 0, K %% 8 == 0");
    IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_tc_selftest: tcgen05 needs an sm_100 device");
    const size_t smem_mn = (size_t)(K / 8) * (32 + N / 4) * 128;
    IMP_REQUIRE(smem_mn <= 200 * 1024, IMP_ERR_ARG, "imp_tc_selftest: tile does not fit shared memory");
    IMP_CUDA(cudaFuncSetAttribute(tc_selftest_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mn));
    tc_selftest_mn_kernel<<<1, 128, smem_mn, (cudaStream_t)stream>>>(d_A, d_B, d_D, N, K, swap_lbo_sbo);
    IMP_LAUNCH_CHECK();
    return 0;
  }
  IMP_REQUIRE((N == 32 || N == 64) && kind >= 0 && kind <= 3, IMP_ERR_ARG, "imp_tc_selftest: N in {32,64}, kind in 0..3");
  if (kind >= 2) {  // A from tensor memory: 2 = bf16, 3 = f16
    IMP_REQUIRE(K > 0 && K % 32 == 0 && K <= 256, IMP_ERR_ARG, "imp_tc_selftest: TS form needs K %% 32 == 0, K <= 256");
    IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_tc_selftest: tcgen05 needs an sm_100 device");
    const size_t smem_ts = (size_t)N * (K / 8) * 16;
    if (kind == 2)
      tc_selftest_ts_kernel<tc::FMT_BF16><<<1, 128, smem_ts, (cudaStream_t)stream>>>(d_A, d_B, d_D, N, K);
    else
      tc_selftest_ts_kernel<tc::FMT_F16><<<1, 128, smem_ts, (cudaStream_t)stream>>>(d_A, d_B, d_D, N, K);
    IMP_LAUNCH_CHECK();
    return 0;
  }
  const int epc = kind == 0 ? 8 : 4;
  IMP_REQUIRE(K > 0 && K % (2 * epc) == 0 && K <= 512, IMP_ERR_ARG, "imp_tc_selftest: K must be a multiple of %d, <= 512", 2 * epc);
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_tc_selftest: tcgen05 needs an sm_100 device");
  const size_t smem = (size_t)(128 + N) * (K / epc) * 16;
  IMP_REQUIRE(smem <= 200 * 1024, IMP_ERR_ARG, "imp_tc_selftest: tile does not fit shared memory");
  if (kind == 0) {
    IMP_CUDA(cudaFuncSetAttribute(tc_selftest_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<0><<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_D, N, K, swap_lbo_sbo);
  } else {
    IMP_CUDA(cudaFuncSetAttribute(tc_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<1><<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_D, N, K, swap_lbo_sbo);
  }
  IMP_LAUNCH_CHECK();
  return 0;
}
