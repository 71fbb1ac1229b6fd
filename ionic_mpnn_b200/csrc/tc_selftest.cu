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

}  // namespace imp

extern "C" int imp_tc_selftest(const float* d_A, const float* d_B, float* d_D, int32_t N, int32_t K, int32_t kind,
                               int32_t swap_lbo_sbo, void* stream) {
  using namespace imp;
  IMP_REQUIRE(d_A && d_B && d_D, IMP_ERR_ARG, "imp_tc_selftest: null pointer");
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
