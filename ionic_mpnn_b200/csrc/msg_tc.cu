// BondMatrixMessage (models/layers.py:100-117) for atom_dim 32 on the tensor cores, from the bond-matrix TABLE (any
// bond_dim: this is the message kernel of the melting-point model, bond_dim = 1024, BASELINE configs[1]) -- the
// north-star's "bond-type-grouped tcgen05 GEMM over gathered h_src rows":
//
//     for every (tower, bond type) bucket b:   M_b = H_src,b . T[b]^T     (E_b x 32) = (E_b x 32) . (32 x 32)
//
// The CSR entries are already grouped by (tower, bond) through bucket_perm (oracle/ref_pack.py, pack_host.cpp).  A CTA
// takes one chunk of <= 128 consecutive slots of ONE bucket: the source rows of its slots are gathered (128 B each, fp32 ->
// 16-bit) into the shared-memory A operand (canonical K-major UMMA layout), the 2 KB operand image of T[b] is copied
// next to it, one elected lane issues two tcgen05.mma (M = 128, N = 32, K = 16) into 32 TMEM columns, and the epilogue
// scales row t by the entry's multiplicity and writes it at the entry's CSR position -- where Reduce
// (models/layers.py:57-83) is the contiguous, deterministic segment sum of imp_segment_sum.
// The kernel is a pure gather/scatter stream (~270 B per entry); the 1,024 FMA per entry that made the SIMT kernels
// L1/FMA-bound (fwd_fp32.cu, profiles/README.md) cost two tensor instructions per 128 entries.  CTAs are small
// (128 threads, 17 KB of shared memory, 32 TMEM columns) so that ~9 of them overlap their gather latencies per SM.
//
// (A first version built agg = Z . Tcat with a block-sparse Z of K = 32 * bond_vocab in shared memory: correct, but the
// A operand traffic of 72 mostly-zero K blocks made it shared-memory-bandwidth-bound at 1.7 ms per step on configs[1].)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

extern "C" int imp_device_is_sm100(void);

namespace imp {
namespace msgtc {

constexpr int D = 32;
constexpr int CHUNK = 128;                   // slots per CTA = MMA M
constexpr int A_BYTES = CHUNK * D * 2;       // 8 KB
constexpr int B_BYTES = D * D * 2;           // 2 KB per bond type

// chunk_ptr[b] = first chunk of bucket b (b over 2 * V_b (tower, bond) buckets); chunk_ptr[2 V_b] = number of chunks
__global__ void chunk_scan_kernel(const int32_t* __restrict__ bucket_ptr, int n_buckets, int32_t* __restrict__ chunk_ptr) {
  if (threadIdx.x == 0) {
    int c = 0;
    for (int b = 0; b < n_buckets; ++b) {
      chunk_ptr[b] = c;
      c += (bucket_ptr[b + 1] - bucket_ptr[b] + CHUNK - 1) / CHUNK;
    }
    chunk_ptr[n_buckets] = c;
  }
}

// Memory access shape: 8 lanes share one 128-byte row (gather of h[src], store of msg[e]), i.e. a warp instruction touches 4
// whole lines instead of 32 partial ones (the one-thread-per-row form spent 8x the L1 wavefronts and ran at 2.1 TB/s).
// The accumulator rows (thread = TMEM lane = slot) are transposed to that shape through a padded shared-memory tile that
// aliases the operand buffers once the MMAs have completed.
constexpr int STG_LD = D + 1;                                   // padded fp32 row
constexpr int SMEM_UNION = CHUNK * STG_LD * 4 > A_BYTES + B_BYTES ? CHUNK * STG_LD * 4 : A_BYTES + B_BYTES;

template <int FMT>
__global__ void __launch_bounds__(CHUNK) grouped_msg_kernel(const int32_t* __restrict__ bucket_ptr, const int32_t* __restrict__ chunk_ptr,
                                                            int n_buckets, int bond_vocab, const int32_t* __restrict__ bucket_perm,
                                                            const int32_t* __restrict__ col_src, const int32_t* __restrict__ edge_bm,
                                                            const float* __restrict__ h, const uint8_t* __restrict__ packed_cat,
                                                            const uint8_t* __restrict__ packed_an, float* __restrict__ msg) {
  __shared__ __align__(128) uint8_t su[SMEM_UNION];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *sA = su, *sB = su + A_BYTES;
  float* stg = reinterpret_cast<float*>(su);
  const int chunk = blockIdx.x;
  if (chunk >= __ldg(chunk_ptr + n_buckets)) return;
  // bucket of this chunk: last b with chunk_ptr[b] <= chunk
  int lo = 0, hi = n_buckets - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
  }
  const int b = lo;
  const int slot0 = __ldg(bucket_ptr + b) + (chunk - __ldg(chunk_ptr + b)) * CHUNK;
  const int n = min(CHUNK, __ldg(bucket_ptr + b + 1) - slot0);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint8_t* tb = (b < bond_vocab ? packed_cat + (int64_t)b * B_BYTES : packed_an + (int64_t)(b - bond_vocab) * B_BYTES);

  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<32>(&tmem_slot);
  // operand image of T[b]: 128 threads x 16 B
  *reinterpret_cast<uint4*>(sB + t * 16) = __ldg(reinterpret_cast<const uint4*>(tb) + t);
  int e = -1, src = 0;
  float mult = 0.f;
  if (t < n) {
    e = __ldg(bucket_perm + slot0 + t);
    mult = (float)((uint32_t)__ldg(edge_bm + e) >> 16);
    src = __ldg(col_src + e);
  }
  // gather: lane group g = lane / 8 takes slot 4 * it + g of this warp, lane % 8 = float4 of the row
  const int g = lane >> 3, q = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + g;  // slot within the warp
    const int rs = __shfl_sync(0xffffffffu, src, r), re = __shfl_sync(0xffffffffu, e, r);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (re >= 0) x = __ldg(reinterpret_cast<const float4*>(h + (int64_t)rs * D) + q);
    // piece p = q / 2 (k in [8p, 8p+8)), half (q % 2) of its 16 bytes
    *reinterpret_cast<uint2*>(sA + (q >> 1) * 2048 + (warp * 32 + r) * 16 + (q & 1) * 8) =
        make_uint2(tc::pack2<FMT>(x.x, x.y), tc::pack2<FMT>(x.z, x.w));
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = tc::make_idesc(FMT, CHUNK, D);
    const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA), 2048, 128), db = tc::make_smem_desc(tc::smem_u32(sB), D * 16, 128);
    if (tc::elect_one()) {
      tc::mma_bf16(tmem, da, db, idesc, false);
      tc::mma_bf16(tmem, da + (uint64_t)(4096 >> 4), db + (uint64_t)((2 * D * 16) >> 4), idesc, true);
      tc::mma_commit(&bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);  // the MMAs have read the operands: the buffers may be reused as the staging tile
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  // (the mbarrier completes only after both MMAs have read the operands, so no further CTA-wide sync is needed here)
#pragma unroll
  for (int c = 0; c < 32; ++c) stg[t * STG_LD + c] = mult * v[c];
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + g;
    const int re = __shfl_sync(0xffffffffu, e, r);
    const float* sr = stg + (warp * 32 + r) * STG_LD + 4 * q;
    const float4 o = make_float4(sr[0], sr[1], sr[2], sr[3]);
    if (re >= 0) reinterpret_cast<float4*>(msg + (int64_t)re * D)[q] = o;
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc<32>(tmem);
  }
}

// 16-bit I/O form of the same GEMM (what the staged tensor forward runs when no intermediates are kept): source rows come from
// the operand-format copy of the atom states (h16 [N, 32], 64-byte rows: a 16-byte lane load IS one K piece of the A operand,
// no conversion) and message rows are written in the operand format too (64 bytes per entry).  Per entry 64 + 64 + 12 bytes
// instead of 128 + 128 + 12.
constexpr int STG16_LD = 17;  // padded row of 16 packed words

template <int FMT>
__global__ void __launch_bounds__(CHUNK) grouped_msg16_kernel(const int32_t* __restrict__ bucket_ptr, const int32_t* __restrict__ chunk_ptr,
                                                              int n_buckets, int bond_vocab, const int32_t* __restrict__ bucket_perm,
                                                              const int32_t* __restrict__ col_src, const int32_t* __restrict__ edge_bm,
                                                              const uint4* __restrict__ h16, const uint8_t* __restrict__ packed_cat,
                                                              const uint8_t* __restrict__ packed_an, uint4* __restrict__ msg16) {
  __shared__ __align__(128) uint8_t su[A_BYTES + B_BYTES];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *sA = su, *sB = su + A_BYTES;
  uint32_t* stg = reinterpret_cast<uint32_t*>(su);  // CHUNK * STG16_LD words <= A_BYTES
  static_assert(CHUNK * STG16_LD * 4 <= A_BYTES + B_BYTES, "staging tile");
  const int chunk = blockIdx.x;
  if (chunk >= __ldg(chunk_ptr + n_buckets)) return;
  int lo = 0, hi = n_buckets - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
  }
  const int b = lo;
  const int slot0 = __ldg(bucket_ptr + b) + (chunk - __ldg(chunk_ptr + b)) * CHUNK;
  const int n = min(CHUNK, __ldg(bucket_ptr + b + 1) - slot0);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint8_t* tb = (b < bond_vocab ? packed_cat + (int64_t)b * B_BYTES : packed_an + (int64_t)(b - bond_vocab) * B_BYTES);
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<32>(&tmem_slot);
  *reinterpret_cast<uint4*>(sB + t * 16) = __ldg(reinterpret_cast<const uint4*>(tb) + t);
  int e = -1, src = 0;
  float mult = 0.f;
  if (t < n) {
    e = __ldg(bucket_perm + slot0 + t);
    mult = (float)((uint32_t)__ldg(edge_bm + e) >> 16);
    src = __ldg(col_src + e);
  }
  // gather: 4 lanes per 64-byte row; lane % 4 = K piece
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = 8 * it + g;
    const int rs = __shfl_sync(0xffffffffu, src, r), re = __shfl_sync(0xffffffffu, e, r);
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (re >= 0) x = __ldg(h16 + (int64_t)rs * 4 + q);
    *reinterpret_cast<uint4*>(sA + q * 2048 + (warp * 32 + r) * 16) = x;
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = tc::make_idesc(FMT, CHUNK, D);
    const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA), 2048, 128), db = tc::make_smem_desc(tc::smem_u32(sB), D * 16, 128);
    if (tc::elect_one()) {
      tc::mma_bf16(tmem, da, db, idesc, false);
      tc::mma_bf16(tmem, da + (uint64_t)(4096 >> 4), db + (uint64_t)((2 * D * 16) >> 4), idesc, true);
      tc::mma_commit(&bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
  for (int c = 0; c < 16; ++c) stg[t * STG16_LD + c] = tc::pack2<FMT>(mult * v[2 * c], mult * v[2 * c + 1]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = 8 * it + g;
    const int re = __shfl_sync(0xffffffffu, e, r);
    const uint32_t* sr = stg + (warp * 32 + r) * STG16_LD + 4 * q;
    if (re >= 0) msg16[(int64_t)re * 4 + q] = make_uint4(sr[0], sr[1], sr[2], sr[3]);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc<32>(tmem);
  }
}

// Pipelined form of grouped_msg16_kernel (the default): persistent CTAs, two operand buffers and two accumulators per CTA.
// While the MMAs and the epilogue of chunk i run, the bucket search, the index loads and the gathers of chunk i + 1 are in
// flight (cp.async 16-byte copies straight into the other A buffer, zero-filled for the slots past the bucket end), so the
// serial search -> indices -> gather -> MMA -> store chain of the one-chunk-per-CTA form is overlapped inside the CTA as well
// as across the CTAs of an SM.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int PIPE_MAX_BUCKETS = 512;  // 2 towers x 256 bond types

struct ChunkIdx {
  int e, src, n;
  float mult;
  const uint8_t* tb;
};

template <int FMT>
__global__ void __launch_bounds__(CHUNK) grouped_msg16_pipe_kernel(const int32_t* __restrict__ bucket_ptr, const int32_t* __restrict__ chunk_ptr,
                                                                   int n_buckets, int bond_vocab, const int32_t* __restrict__ bucket_perm,
                                                                   const int32_t* __restrict__ col_src, const int32_t* __restrict__ edge_bm,
                                                                   const uint4* __restrict__ h16, const uint8_t* __restrict__ packed_cat,
                                                                   const uint8_t* __restrict__ packed_an, uint4* __restrict__ msg16) {
  __shared__ __align__(128) uint8_t su[2][A_BYTES + B_BYTES];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_cptr[PIPE_MAX_BUCKETS + 1], s_bptr[PIPE_MAX_BUCKETS + 1];  // chunk / slot offsets of the buckets
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n_chunks = __ldg(chunk_ptr + n_buckets);
  if ((int)blockIdx.x >= n_chunks) return;
  for (int i = t; i <= n_buckets; i += CHUNK) s_cptr[i] = __ldg(chunk_ptr + i), s_bptr[i] = __ldg(bucket_ptr + i);
  int cur_bucket = 0;  // a CTA walks its chunks in increasing order: the bucket index only moves forward
  if (t == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
  const int g = lane >> 2, q = lane & 3;
  const uint32_t idesc = tc::make_idesc(FMT, CHUNK, D);

  auto locate = [&](int chunk) {  // bucket search + this thread's slot
    ChunkIdx c;
    while (s_cptr[cur_bucket + 1] <= chunk) ++cur_bucket;  // shared-memory copies: no global latency in the search
    const int b = cur_bucket;
    const int slot0 = s_bptr[b] + (chunk - s_cptr[b]) * CHUNK;
    c.n = min(CHUNK, s_bptr[b + 1] - slot0);
    c.tb = b < bond_vocab ? packed_cat + (int64_t)b * B_BYTES : packed_an + (int64_t)(b - bond_vocab) * B_BYTES;
    c.e = -1, c.src = 0, c.mult = 0.f;
    if (t < c.n) {
      c.e = __ldg(bucket_perm + slot0 + t);
      c.mult = (float)((uint32_t)__ldg(edge_bm + c.e) >> 16);
      c.src = __ldg(col_src + c.e);
    }
    return c;
  };
  auto gather = [&](const ChunkIdx& c, int p) {  // asynchronous: A rows (4 lanes per 64-byte row) + the T[b] image
    uint8_t *sA = su[p], *sB = su[p] + A_BYTES;
    cp_async16(sB + t * 16, c.tb + t * 16, true);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = 8 * it + g;
      const int rs = __shfl_sync(0xffffffffu, c.src, r), re = __shfl_sync(0xffffffffu, c.e, r);
      cp_async16(sA + q * 2048 + (warp * 32 + r) * 16, h16 + (int64_t)rs * 4 + q, re >= 0);
    }
    cp_async_commit();
  };

  tc::fence_before_thread_sync();
  __syncthreads();  // bucket offsets, barriers, TMEM address
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  ChunkIdx cur = locate(blockIdx.x);
  gather(cur, 0);
  int i = 0;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++i) {
    const int p = i & 1;
    const bool has_next = chunk + (int)gridDim.x < n_chunks;
    ChunkIdx nxt = cur;
    if (has_next) {
      nxt = locate(chunk + gridDim.x);
      gather(nxt, p ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_thread_sync();
      const uint32_t sa = tc::smem_u32(su[p]);
      const uint64_t da = tc::make_smem_desc(sa, 2048, 128), db = tc::make_smem_desc(sa + A_BYTES, D * 16, 128);
      if (tc::elect_one()) {
        tc::mma_bf16(tmem + p * D, da, db, idesc, false);
        tc::mma_bf16(tmem + p * D, da + (uint64_t)(4096 >> 4), db + (uint64_t)((2 * D * 16) >> 4), idesc, true);
        tc::mma_commit(&bar[p]);
      }
      __syncwarp();
    }
    tc::mbar_wait(&bar[p], (uint32_t)((i >> 1) & 1));
    tc::fence_after_thread_sync();
    float v[32];
    tc::tmem_ld32(tmem + p * D + ((uint32_t)(warp * 32) << 16), v);
    uint32_t* stg = reinterpret_cast<uint32_t*>(su[p]);  // the operands of this chunk have been consumed
#pragma unroll
    for (int c = 0; c < 16; ++c) stg[t * STG16_LD + c] = tc::pack2<FMT>(cur.mult * v[2 * c], cur.mult * v[2 * c + 1]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = 8 * it + g;
      const int re = __shfl_sync(0xffffffffu, cur.e, r);
      const uint32_t* sr = stg + (warp * 32 + r) * STG16_LD + 4 * q;
      if (re >= 0) msg16[(int64_t)re * 4 + q] = make_uint4(sr[0], sr[1], sr[2], sr[3]);
    }
    tc::fence_before_thread_sync();
    __syncthreads();  // buffer p and accumulator p are free for chunk i + 2
    tc::fence_after_thread_sync();
    cur = nxt;
  }
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

// Planned form (what the model's forward runs): the per-batch index work is done ONCE per batch instead of once per step and
// per chunk -- imp_edge_messages_tc16_plan writes the chunk offsets and bucket-ordered copies of the source atoms and of
// bond | multiplicity, so that the per-step kernel reads every index with independent coalesced loads (no perm -> src chain,
// no warp shuffles: ncu showed 13 long-scoreboard + 12 mio-throttle stalls per issue on exactly those) two chunks ahead
// of their use.  Same persistent double-buffered structure as grouped_msg16_pipe_kernel, bit-identical rows.
struct PlanIdx {
  int gsrc[4], gpos[4];
  float mult;
  const uint8_t* tb;
};

__global__ void plan_kernel(const int32_t* __restrict__ bucket_perm, const int32_t* __restrict__ col_src, const int32_t* __restrict__ edge_bm,
                            int n_unique, int32_t* __restrict__ bsrc, int32_t* __restrict__ bbm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_unique) return;
  const int e = __ldg(bucket_perm + i);
  bsrc[i] = __ldg(col_src + e);
  bbm[i] = __ldg(edge_bm + e);
}

template <int FMT>
__global__ void __launch_bounds__(CHUNK) grouped_msg16_planned_kernel(const int32_t* __restrict__ bucket_ptr, const int32_t* __restrict__ chunk_ptr,
                                                                      int n_buckets, int bond_vocab, const int32_t* __restrict__ bucket_perm,
                                                                      const int32_t* __restrict__ bsrc, const int32_t* __restrict__ bbm,
                                                                      const uint4* __restrict__ h16, const uint8_t* __restrict__ packed_cat,
                                                                      const uint8_t* __restrict__ packed_an, uint4* __restrict__ msg16) {
  __shared__ __align__(128) uint8_t su[2][A_BYTES + B_BYTES];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_cptr[PIPE_MAX_BUCKETS + 1], s_bptr[PIPE_MAX_BUCKETS + 1];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n_chunks = __ldg(chunk_ptr + n_buckets);
  if ((int)blockIdx.x >= n_chunks) return;
  for (int i = t; i <= n_buckets; i += CHUNK) s_cptr[i] = __ldg(chunk_ptr + i), s_bptr[i] = __ldg(bucket_ptr + i);
  int cur_bucket = 0;
  if (t == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
  const int g = lane >> 2, q = lane & 3;
  const uint32_t idesc = tc::make_idesc(FMT, CHUNK, D);

  auto load_idx = [&](int chunk) {  // independent coalesced loads; chunks are visited in increasing order
    PlanIdx c;
    while (s_cptr[cur_bucket + 1] <= chunk) ++cur_bucket;
    const int b = cur_bucket;
    const int slot0 = s_bptr[b] + (chunk - s_cptr[b]) * CHUNK;
    const int n = min(CHUNK, s_bptr[b + 1] - slot0);
    c.tb = b < bond_vocab ? packed_cat + (int64_t)b * B_BYTES : packed_an + (int64_t)(b - bond_vocab) * B_BYTES;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = warp * 32 + 8 * it + g;
      const bool ok = r < n;
      c.gsrc[it] = ok ? __ldg(bsrc + slot0 + r) : -1;
      c.gpos[it] = ok ? __ldg(bucket_perm + slot0 + r) : -1;
    }
    c.mult = t < n ? (float)((uint32_t)__ldg(bbm + slot0 + t) >> 16) : 0.f;
    return c;
  };
  auto gather = [&](const PlanIdx& c, int p) {
    uint8_t *sA = su[p], *sB = su[p] + A_BYTES;
    cp_async16(sB + t * 16, c.tb + t * 16, true);
#pragma unroll
    for (int it = 0; it < 4; ++it)
      cp_async16(sA + q * 2048 + (warp * 32 + 8 * it + g) * 16, h16 + (int64_t)max(c.gsrc[it], 0) * 4 + q, c.gsrc[it] >= 0);
    cp_async_commit();
  };

  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  const int G = gridDim.x;
  PlanIdx i0 = load_idx(blockIdx.x), i1 = i0, i2 = i0;
  if ((int)blockIdx.x + G < n_chunks) i1 = load_idx(blockIdx.x + G);
  gather(i0, 0);
  int i = 0;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += G, ++i) {
    const int p = i & 1;
    const bool has_next = chunk + G < n_chunks;
    if (has_next) gather(i1, p ^ 1);                               // indices loaded an iteration ago
    if (chunk + 2 * G < n_chunks) i2 = load_idx(chunk + 2 * G);    // in flight until the next iteration
    if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_thread_sync();
      const uint32_t sa = tc::smem_u32(su[p]);
      const uint64_t da = tc::make_smem_desc(sa, 2048, 128), db = tc::make_smem_desc(sa + A_BYTES, D * 16, 128);
      if (tc::elect_one()) {
        tc::mma_bf16(tmem + p * D, da, db, idesc, false);
        tc::mma_bf16(tmem + p * D, da + (uint64_t)(4096 >> 4), db + (uint64_t)((2 * D * 16) >> 4), idesc, true);
        tc::mma_commit(&bar[p]);
      }
      __syncwarp();
    }
    tc::mbar_wait(&bar[p], (uint32_t)((i >> 1) & 1));
    tc::fence_after_thread_sync();
    float v[32];
    tc::tmem_ld32(tmem + p * D + ((uint32_t)(warp * 32) << 16), v);
    uint32_t* stg = reinterpret_cast<uint32_t*>(su[p]);
#pragma unroll
    for (int c = 0; c < 16; ++c) stg[t * STG16_LD + c] = tc::pack2<FMT>(i0.mult * v[2 * c], i0.mult * v[2 * c + 1]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint32_t* sr = stg + (warp * 32 + 8 * it + g) * STG16_LD + 4 * q;
      if (i0.gpos[it] >= 0) msg16[(int64_t)i0.gpos[it] * 4 + q] = make_uint4(sr[0], sr[1], sr[2], sr[3]);
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    i0 = i1, i1 = i2;
  }
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

// Embedding(atom) (train_viscosity.py:163,171) writing the fp32 state and its operand-format copy
template <int FMT>
__global__ void embed16_kernel(const float4* __restrict__ emb, const int* __restrict__ atom_id, int64_t total4, int atom_vocab,
                               float4* __restrict__ out, uint2* __restrict__ out16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int64_t v = i / (D / 4);
  const int c = (int)(i - v * (D / 4));
  const int id = min(max(__ldg(atom_id + v), 0), atom_vocab - 1);
  const float4 x = __ldg(emb + (int64_t)id * (D / 4) + c);
  out[i] = x;
  out16[i] = make_uint2(tc::pack2<FMT>(x.x, x.y), tc::pack2<FMT>(x.z, x.w));
}

// table [V_b, d, d] fp32 (T[b][l][m], models/layers.py:108-112) -> per bond type one 2 KB image of the canonical K-major
// B operand [piece c of 4][n = l of 32][8 halfs]: B[n][kk] = T[b][n][kk], kk = c*8 + i = m.
template <int FMT>
__global__ void msg_pack_kernel(const float* __restrict__ table, int bond_vocab, uint16_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= bond_vocab * D * D) return;
  const int i = idx & 7, n = (idx >> 3) & 31, c = (idx >> 8) & 3, b = idx >> 10;
  out[idx] = tc::cvt16<FMT>(table[((int64_t)b * D + n) * D + c * 8 + i]);
}

}  // namespace msgtc
}  // namespace imp

using namespace imp;

extern "C" int64_t imp_message_pack_bytes(int32_t bond_vocab, int32_t d) {
  if (d != msgtc::D || bond_vocab <= 0) return IMP_ERR_DIM;
  return (int64_t)bond_vocab * msgtc::B_BYTES;
}

extern "C" int imp_message_pack(const float* d_table, int32_t bond_vocab, int32_t d, int32_t flags, void* d_packed, void* stream) {
  IMP_REQUIRE(d == msgtc::D, IMP_ERR_DIM, "imp_message_pack: atom_dim %d not supported by the tensor path (32)", d);
  IMP_REQUIRE(d_table && d_packed && bond_vocab > 0 && bond_vocab <= 65535, IMP_ERR_ARG, "imp_message_pack: bad arguments");
  const int n = bond_vocab * msgtc::D * msgtc::D;
  if (flags & IMP_TC_FP16)
    msgtc::msg_pack_kernel<tc::FMT_F16><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_table, bond_vocab, (uint16_t*)d_packed);
  else
    msgtc::msg_pack_kernel<tc::FMT_BF16><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_table, bond_vocab, (uint16_t*)d_packed);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t imp_edge_messages_tc_workspace_bytes(int32_t bond_vocab) { return (int64_t)(2 * bond_vocab + 1) * 4; }

extern "C" int imp_edge_messages_tc(const imp_graph_t* g, const float* d_h, int32_t d, const void* d_packed_cat,
                                    const void* d_packed_an, int32_t flags, float* d_msg, void* d_workspace, void* stream) {
  IMP_REQUIRE(g, IMP_ERR_ARG, "imp_edge_messages_tc: graph is null");
  IMP_REQUIRE(d == msgtc::D, IMP_ERR_DIM, "imp_edge_messages_tc: atom_dim %d not supported by the tensor path (32)", d);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(d_h && d_msg && d_packed_cat && d_packed_an && d_workspace && g->bucket_ptr && g->bucket_perm && g->col_src && g->edge_bm,
              IMP_ERR_ARG, "imp_edge_messages_tc: null pointer (the bond-bucket permutation is required)");
  IMP_REQUIRE(g->bond_vocab > 0 && g->bond_vocab <= 65535, IMP_ERR_ARG, "imp_edge_messages_tc: bond vocabulary out of range");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_edge_messages_tc: tcgen05 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * g->bond_vocab;
  int32_t* chunk_ptr = reinterpret_cast<int32_t*>(d_workspace);
  msgtc::chunk_scan_kernel<<<1, 32, 0, st>>>(g->bucket_ptr, nb, chunk_ptr);
  IMP_LAUNCH_CHECK();
  const unsigned grid = (unsigned)(ceil_div(g->n_unique, msgtc::CHUNK) + nb);  // upper bound; surplus CTAs exit at once
  const uint8_t *pc = reinterpret_cast<const uint8_t*>(d_packed_cat), *pa = reinterpret_cast<const uint8_t*>(d_packed_an);
  if (flags & IMP_TC_FP16)
    msgtc::grouped_msg_kernel<tc::FMT_F16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                         g->col_src, g->edge_bm, d_h, pc, pa, d_msg);
  else
    msgtc::grouped_msg_kernel<tc::FMT_BF16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                          g->col_src, g->edge_bm, d_h, pc, pa, d_msg);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_embed_atoms16(const float* d_atom_emb, int32_t atom_vocab, const int32_t* d_atom_id, int32_t n_atoms, int32_t d,
                                 int32_t flags, float* d_h0, void* d_h0_16, void* stream) {
  IMP_REQUIRE(d == msgtc::D, IMP_ERR_DIM, "imp_embed_atoms16: atom_dim %d not supported by the tensor path (32)", d);
  IMP_REQUIRE(n_atoms >= 0 && atom_vocab > 0, IMP_ERR_ARG, "imp_embed_atoms16: bad sizes");
  if (n_atoms == 0) return 0;
  IMP_REQUIRE(d_atom_emb && d_atom_id && d_h0 && d_h0_16, IMP_ERR_ARG, "imp_embed_atoms16: null pointer");
  const int64_t total4 = (int64_t)n_atoms * (d / 4);
  const unsigned blocks = (unsigned)ceil_div(total4, 256);
  if (flags & IMP_TC_FP16)
    msgtc::embed16_kernel<tc::FMT_F16><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_atom_emb), d_atom_id, total4,
                                                                               atom_vocab, reinterpret_cast<float4*>(d_h0),
                                                                               reinterpret_cast<uint2*>(d_h0_16));
  else
    msgtc::embed16_kernel<tc::FMT_BF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_atom_emb), d_atom_id, total4,
                                                                                atom_vocab, reinterpret_cast<float4*>(d_h0),
                                                                                reinterpret_cast<uint2*>(d_h0_16));
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_edge_messages_tc16(const imp_graph_t* g, const void* d_h16, int32_t d, const void* d_packed_cat,
                                      const void* d_packed_an, int32_t flags, void* d_msg16, void* d_workspace, void* stream) {
  IMP_REQUIRE(g, IMP_ERR_ARG, "imp_edge_messages_tc16: graph is null");
  IMP_REQUIRE(d == msgtc::D, IMP_ERR_DIM, "imp_edge_messages_tc16: atom_dim %d not supported by the tensor path (32)", d);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(d_h16 && d_msg16 && d_packed_cat && d_packed_an && d_workspace && g->bucket_ptr && g->bucket_perm && g->col_src && g->edge_bm,
              IMP_ERR_ARG, "imp_edge_messages_tc16: null pointer (the bond-bucket permutation is required)");
  IMP_REQUIRE(g->bond_vocab > 0 && g->bond_vocab <= 65535, IMP_ERR_ARG, "imp_edge_messages_tc16: bond vocabulary out of range");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_edge_messages_tc16: tcgen05 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * g->bond_vocab;
  int32_t* chunk_ptr = reinterpret_cast<int32_t*>(d_workspace);
  msgtc::chunk_scan_kernel<<<1, 32, 0, st>>>(g->bucket_ptr, nb, chunk_ptr);
  IMP_LAUNCH_CHECK();
  const unsigned grid_all = (unsigned)(ceil_div(g->n_unique, msgtc::CHUNK) + nb);  // upper bound on the number of chunks
  const uint8_t *pc = reinterpret_cast<const uint8_t*>(d_packed_cat), *pa = reinterpret_cast<const uint8_t*>(d_packed_an);
  if (!(flags & IMP_TC_MSG_ONE_CHUNK_PER_CTA) && nb <= msgtc::PIPE_MAX_BUCKETS) {  // default: persistent, software-pipelined CTAs
    int dev = 0, sms = 148;
    IMP_CUDA(cudaGetDevice(&dev));
    IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = grid_all < (unsigned)(8 * sms) ? grid_all : (unsigned)(8 * sms);
    if (flags & IMP_TC_FP16)
      msgtc::grouped_msg16_pipe_kernel<tc::FMT_F16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                                  g->col_src, g->edge_bm, reinterpret_cast<const uint4*>(d_h16),
                                                                                  pc, pa, reinterpret_cast<uint4*>(d_msg16));
    else
      msgtc::grouped_msg16_pipe_kernel<tc::FMT_BF16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                                   g->col_src, g->edge_bm, reinterpret_cast<const uint4*>(d_h16),
                                                                                   pc, pa, reinterpret_cast<uint4*>(d_msg16));
    IMP_LAUNCH_CHECK();
    return 0;
  }
  const unsigned grid = grid_all;
  if (flags & IMP_TC_FP16)
    msgtc::grouped_msg16_kernel<tc::FMT_F16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                           g->col_src, g->edge_bm, reinterpret_cast<const uint4*>(d_h16), pc,
                                                                           pa, reinterpret_cast<uint4*>(d_msg16));
  else
    msgtc::grouped_msg16_kernel<tc::FMT_BF16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                            g->col_src, g->edge_bm, reinterpret_cast<const uint4*>(d_h16), pc,
                                                                            pa, reinterpret_cast<uint4*>(d_msg16));
  IMP_LAUNCH_CHECK();
  return 0;
}

// plan = [chunk_ptr: 2 V_b + 1 ints, padded to 1024 ints][bsrc: Eu ints][bbm: Eu ints]
extern "C" int64_t imp_edge_messages_tc16_plan_bytes(int32_t n_unique, int32_t bond_vocab) {
  if (n_unique < 0 || bond_vocab <= 0 || 2 * bond_vocab > msgtc::PIPE_MAX_BUCKETS) return IMP_ERR_ARG;
  return (int64_t)4 * (1024 + 2 * (int64_t)n_unique);
}

extern "C" int imp_edge_messages_tc16_plan(const imp_graph_t* g, void* d_plan, void* stream) {
  IMP_REQUIRE(g && d_plan, IMP_ERR_ARG, "imp_edge_messages_tc16_plan: null pointer");
  IMP_REQUIRE(g->bond_vocab > 0 && 2 * g->bond_vocab <= msgtc::PIPE_MAX_BUCKETS, IMP_ERR_ARG,
              "imp_edge_messages_tc16_plan: bond vocabulary %d out of range (1..256)", g->bond_vocab);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(g->bucket_ptr && g->bucket_perm && g->col_src && g->edge_bm, IMP_ERR_ARG,
              "imp_edge_messages_tc16_plan: the bond-bucket permutation is required");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* plan = reinterpret_cast<int32_t*>(d_plan);
  msgtc::chunk_scan_kernel<<<1, 32, 0, st>>>(g->bucket_ptr, 2 * g->bond_vocab, plan);
  IMP_LAUNCH_CHECK();
  msgtc::plan_kernel<<<(unsigned)ceil_div(g->n_unique, 256), 256, 0, st>>>(g->bucket_perm, g->col_src, g->edge_bm, g->n_unique, plan + 1024,
                                                                           plan + 1024 + g->n_unique);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_edge_messages_tc16_planned(const imp_graph_t* g, const void* d_plan, const void* d_h16, int32_t d,
                                              const void* d_packed_cat, const void* d_packed_an, int32_t flags, void* d_msg16,
                                              void* stream) {
  IMP_REQUIRE(g, IMP_ERR_ARG, "imp_edge_messages_tc16_planned: graph is null");
  IMP_REQUIRE(d == msgtc::D, IMP_ERR_DIM, "imp_edge_messages_tc16_planned: atom_dim %d not supported by the tensor path (32)", d);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(d_plan && d_h16 && d_msg16 && d_packed_cat && d_packed_an && g->bucket_ptr && g->bucket_perm, IMP_ERR_ARG,
              "imp_edge_messages_tc16_planned: null pointer");
  IMP_REQUIRE(g->bond_vocab > 0 && 2 * g->bond_vocab <= msgtc::PIPE_MAX_BUCKETS, IMP_ERR_ARG,
              "imp_edge_messages_tc16_planned: bond vocabulary out of range");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_edge_messages_tc16_planned: tcgen05 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * g->bond_vocab;
  const int32_t* plan = reinterpret_cast<const int32_t*>(d_plan);
  int dev = 0, sms = 148;
  IMP_CUDA(cudaGetDevice(&dev));
  IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid_all = (unsigned)(ceil_div(g->n_unique, msgtc::CHUNK) + nb);
  const unsigned grid = grid_all < (unsigned)(8 * sms) ? grid_all : (unsigned)(8 * sms);
  const uint8_t *pc = reinterpret_cast<const uint8_t*>(d_packed_cat), *pa = reinterpret_cast<const uint8_t*>(d_packed_an);
  if (flags & IMP_TC_FP16)
    msgtc::grouped_msg16_planned_kernel<tc::FMT_F16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, plan, nb, g->bond_vocab, g->bucket_perm,
                                                                                   plan + 1024, plan + 1024 + g->n_unique,
                                                                                   reinterpret_cast<const uint4*>(d_h16), pc, pa,
                                                                                   reinterpret_cast<uint4*>(d_msg16));
  else
    msgtc::grouped_msg16_planned_kernel<tc::FMT_BF16><<<grid, msgtc::CHUNK, 0, st>>>(g->bucket_ptr, plan, nb, g->bond_vocab, g->bucket_perm,
                                                                                    plan + 1024, plan + 1024 + g->n_unique,
                                                                                    reinterpret_cast<const uint4*>(d_h16), pc, pa,
                                                                                    reinterpret_cast<uint4*>(d_msg16));
  IMP_LAUNCH_CHECK();
  return 0;
}
