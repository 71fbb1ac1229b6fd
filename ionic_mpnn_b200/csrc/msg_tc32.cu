// BondMatrixMessage (models/layers.py:100-117) and its transpose for atom_dim 32 on the tensor cores with fp32-class accuracy:
// the bucket-grouped message kernel of the TRAINING step (train_viscosity.py:227-230: forward, and the backward with respect
// to the atom states) and of the fp32 layer API.
//
//     for every (tower, bond type) bucket b:   M_b = X_src,b . T[b]^T   (forward)      or   X_src,b . T[b]   (transposed)
//
// The fp32 SIMT kernel (grouped_msg_f32_kernel, fwd_fp32.cu) spends 1,024 FFMA and 256 shared-memory weight reads per entry
// and runs at a third of the FMA pipe; here a CTA takes one chunk of <= 128 consecutive slots of ONE bucket, gathers the
// source rows (128 B each) into two shared-memory A operands -- x = hi + lo, hi = x with 13 low significand bits cleared,
// lo = x - hi (exact) -- stages T[b] (or its transpose) split the same way, and one elected lane issues twelve
// tcgen05.mma kind::tf32 (M = 128, N = 32, K = 8: hi.hi + hi.lo + lo.hi per K step; ~2^-21 relative per product, as
// csrc/bwd_tc.cu) into 32 TMEM columns.  The epilogue scales row t by the entry's multiplicity and writes it at the entry's CSR
// position.  Per entry 128 + 128 + 12 bytes: a gather / scatter stream.  Deterministic.
#include "common.cuh"
#include "tc_common.cuh"

extern "C" int imp_device_is_sm100(void);

namespace imp {
namespace msg32 {

constexpr int D = 32;
constexpr int CHUNK = 128;                  // slots per CTA = MMA M
constexpr int A_LBO = (CHUNK + 1) * 16;     // bytes between K chunks (4 tf32) of the A operands: padded by one row so that the 8
                                            // lanes that share a source row store to 8 different bank groups
constexpr int A_BYTES = (D / 4) * A_LBO;    // one A operand (hi or lo)
constexpr int STG_LD = D + 1;               // padded fp32 row of the output staging tile (aliases the A operands)
static_assert(CHUNK * STG_LD * 4 <= 2 * A_BYTES, "staging tile fits the A operands");

struct Smem {
  unsigned char a[2][A_BYTES];
  float b[2][D * D];
  uint64_t bar;
  uint32_t tmem_slot;
};

// hi = x with its 13 low significand bits cleared -- what kind::tf32 reads of a raw fp32 operand --, lo = x - hi (exact in fp32)
// ROUNDED to tf32 (the MMA would truncate it: biased, and twice the error)
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  uint32_t l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}
// both terms rounded (operands that are staged explicitly: the bond matrices)
__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}
// planned kernel: RN selects the rounded forms above (fp32 inference route), else plain truncation splits (training: 4 % faster)
template <bool RN>
__device__ __forceinline__ void split_a(float x, float& hi, float& lo) {
  if (RN) {
    split(x, hi, lo);
  } else {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
  }
}
template <bool RN>
__device__ __forceinline__ void split_b(float x, float& hi, float& lo) {
  if (RN) split_rn(x, hi, lo);
  else split_a<false>(x, hi, lo);
}

template <bool TRANSPOSED>
__global__ void __launch_bounds__(CHUNK) grouped_msg_tf32x3_kernel(const int32_t* __restrict__ bucket_ptr, const int32_t* __restrict__ chunk_ptr,
                                                                   int n_buckets, int bond_vocab, const int32_t* __restrict__ bucket_perm,
                                                                   const int32_t* __restrict__ col_src, const int32_t* __restrict__ edge_bm,
                                                                   const float* __restrict__ x, const float* __restrict__ tab_cat,
                                                                   const float* __restrict__ tab_an, float* __restrict__ msg) {
  __shared__ __align__(128) Smem s;
  const int chunk = blockIdx.x;
  if (chunk >= __ldg(chunk_ptr + n_buckets)) return;
  int lo = 0, hi = n_buckets - 1;
  while (lo < hi) {  // bucket of this chunk: last b with chunk_ptr[b] <= chunk
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid - 1;
  }
  const int b = lo;
  const int slot0 = __ldg(bucket_ptr + b) + (chunk - __ldg(chunk_ptr + b)) * CHUNK;
  const int n = min(CHUNK, __ldg(bucket_ptr + b + 1) - slot0);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float* tb = b < bond_vocab ? tab_cat + (int64_t)b * D * D : tab_an + (int64_t)(b - bond_vocab) * D * D;

  if (t == 0) {
    tc::mbar_init(&s.bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<32>(&s.tmem_slot);
  // B[n][k] (K-major, element (n, k) at chunk_off(n, k / 4, 32) + (k % 4) * 4): forward  out[l] = sum_m T[l][m] x[m]: n = l, k = m;
  // transposed  out[m] = sum_l T[l][m] x[l]: n = m, k = l.  T[b] is row-major [l][m].
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f4 = t + CHUNK * i;             // float4 index into T[b]: row l = f4 / 8, columns 4 (f4 % 8) ..
    const float4 v = __ldg(reinterpret_cast<const float4*>(tb) + f4);
    const int l = f4 >> 3, m0 = (f4 & 7) * 4;
    const float vv[4] = {v.x, v.y, v.z, v.w};
    if (!TRANSPOSED) {  // (n, k) = (l, m0 + j): one 16-byte chunk
      float h4[4], l4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_rn(vv[j], h4[j], l4[j]);
      const int o = tc::chunk_off(l, m0 / 4, D) / 4;
      *reinterpret_cast<float4*>(&s.b[0][o]) = make_float4(h4[0], h4[1], h4[2], h4[3]);
      *reinterpret_cast<float4*>(&s.b[1][o]) = make_float4(l4[0], l4[1], l4[2], l4[3]);
    } else {  // (n, k) = (m0 + j, l)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a, c;
        split_rn(vv[j], a, c);
        const int o = (tc::chunk_off(m0 + j, l / 4, D) + (l % 4) * 4) / 4;
        s.b[0][o] = a, s.b[1][o] = c;
      }
    }
  }
  int e = -1, src = 0;
  float mult = 0.f;
  if (t < n) {
    e = __ldg(bucket_perm + slot0 + t);
    mult = (float)((uint32_t)__ldg(edge_bm + e) >> 16);
    src = __ldg(col_src + e);
  }
  // gather: lane group g = lane / 8 takes slot 4 * it + g of this warp, lane % 8 = float4 of the row = K chunk q
  const int g = lane >> 3, q = lane & 7;
  float4 xr[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {  // all eight loads in flight before the first is used
    const int r = 4 * it + g;
    const int rs = __shfl_sync(0xffffffffu, src, r), re = __shfl_sync(0xffffffffu, e, r);
    xr[it] = re >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (int64_t)rs * D) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + g;
    float4 h4, l4;
    split(xr[it].x, h4.x, l4.x), split(xr[it].y, h4.y, l4.y), split(xr[it].z, h4.z, l4.z), split(xr[it].w, h4.w, l4.w);
    const int o = q * A_LBO + (warp * 32 + r) * 16;
    *reinterpret_cast<float4*>(s.a[0] + o) = h4;
    *reinterpret_cast<float4*>(s.a[1] + o) = l4;
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = s.tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, CHUNK, D);
    const uint64_t dah = tc::make_smem_desc(tc::smem_u32(s.a[0]), A_LBO, 128), dal = tc::make_smem_desc(tc::smem_u32(s.a[1]), A_LBO, 128);
    const uint64_t dbh = tc::make_smem_desc(tc::smem_u32(s.b[0]), D * 16, 128), dbl = tc::make_smem_desc(tc::smem_u32(s.b[1]), D * 16, 128);
    if (tc::elect_one()) {
#pragma unroll
      for (int ks = 0; ks < D / 8; ++ks) {  // K = 8 per MMA = two 16-byte chunks
        const uint64_t ao = (uint64_t)((ks * 2 * A_LBO) >> 4), bo = (uint64_t)((ks * 2 * D * 16) >> 4);
        tc::mma_tf32(tmem, dah + ao, dbh + bo, idesc, ks > 0);
        tc::mma_tf32(tmem, dah + ao, dbl + bo, idesc, true);
        tc::mma_tf32(tmem, dal + ao, dbh + bo, idesc, true);
      }
      tc::mma_commit(&s.bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&s.bar, 0);  // the MMAs have read the operands: the A buffers may be reused as the staging tile
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  float* stg = reinterpret_cast<float*>(s.a);
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

// ---------------------------------------------------------------------------------------------------------------------------
// Planned, persistent form (what the training step runs): the per-batch index plan of csrc/msg_tc.cu (imp_edge_messages_tc16_plan:
// chunk offsets, bucket-ordered source atoms and bond | multiplicity) is read with independent coalesced loads two chunks
// ahead; the source rows of chunk i + 1 arrive by cp.async (16-byte copies straight into the other A buffer, zero-filled past
// the bucket end) while chunk i's MMAs and epilogue run.  The copied rows ARE the hi operand -- kind::tf32 ignores the 13 low
// significand bits of its inputs, which is the truncation split(x).hi -- so only lo = x - trunc(x) is computed, from shared
// memory.  Two accumulators (64 TMEM columns), 58 KB of shared memory: three CTAs per SM.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int MAX_BUCKETS = 512;  // 2 towers x 256 bond types
constexpr int RAW_BYTES = CHUNK * STG_LD * 4 > A_BYTES ? CHUNK * STG_LD * 4 : A_BYTES;  // a raw A buffer doubles as the output staging tile

struct PSmem {
  unsigned char raw[2][RAW_BYTES];  // gathered fp32 rows (= the hi operand), double-buffered
  unsigned char lo[A_BYTES];
  float b[2][D * D];
  uint64_t bar[2];
  uint32_t tmem_slot;
  int cptr[MAX_BUCKETS + 1], bptr[MAX_BUCKETS + 1];
};

struct PIdx {
  int gsrc[8], gpos[8];
  float mult;
  const float* tb;
};

template <bool TRANSPOSED, bool RN>
__global__ void __launch_bounds__(CHUNK) grouped_msg_tf32x3_planned_kernel(const int32_t* __restrict__ bucket_ptr, const int32_t* __restrict__ chunk_ptr,
                                                                           int n_buckets, int bond_vocab, const int32_t* __restrict__ bucket_perm,
                                                                           const int32_t* __restrict__ bsrc, const int32_t* __restrict__ bbm,
                                                                           const float* __restrict__ x, const float* __restrict__ tab_cat,
                                                                           const float* __restrict__ tab_an, float* __restrict__ msg) {
  extern __shared__ __align__(128) unsigned char psm_raw[];
  PSmem& s = *reinterpret_cast<PSmem*>(psm_raw);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n_chunks = __ldg(chunk_ptr + n_buckets);
  if ((int)blockIdx.x >= n_chunks) return;
  for (int i = t; i <= n_buckets; i += CHUNK) s.cptr[i] = __ldg(chunk_ptr + i), s.bptr[i] = __ldg(bucket_ptr + i);
  int cur_bucket = 0;
  if (t == 0) {
    tc::mbar_init(&s.bar[0], 1);
    tc::mbar_init(&s.bar[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<64>(&s.tmem_slot);
  const int g = lane >> 3, q = lane & 7;  // lane group g takes slot 4 it + g of this warp, q = float4 of the row = K chunk
  const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, CHUNK, D);

  auto load_idx = [&](int chunk) {  // independent coalesced loads; chunks are visited in increasing order
    PIdx c;
    while (s.cptr[cur_bucket + 1] <= chunk) ++cur_bucket;
    const int b = cur_bucket;
    const int slot0 = s.bptr[b] + (chunk - s.cptr[b]) * CHUNK;
    const int n = min(CHUNK, s.bptr[b + 1] - slot0);
    c.tb = b < bond_vocab ? tab_cat + (int64_t)b * D * D : tab_an + (int64_t)(b - bond_vocab) * D * D;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = warp * 32 + 4 * it + g;
      const bool ok = r < n;
      c.gsrc[it] = ok ? __ldg(bsrc + slot0 + r) : -1;
      c.gpos[it] = ok ? __ldg(bucket_perm + slot0 + r) : -1;
    }
    c.mult = t < n ? (float)((uint32_t)__ldg(bbm + slot0 + t) >> 16) : 0.f;
    return c;
  };
  auto gather = [&](const PIdx& c, int p) {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      cp_async16(s.raw[p] + q * A_LBO + (warp * 32 + 4 * it + g) * 16, x + (int64_t)max(c.gsrc[it], 0) * D + 4 * q, c.gsrc[it] >= 0);
    cp_async_commit();
  };

  __syncthreads();  // cptr / bptr
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = s.tmem_slot;
  const int G = gridDim.x;
  PIdx i0 = load_idx(blockIdx.x), i1 = i0, i2 = i0;
  if ((int)blockIdx.x + G < n_chunks) i1 = load_idx(blockIdx.x + G);
  gather(i0, 0);
  int i = 0;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += G, ++i) {
    const int p = i & 1;
    const bool has_next = chunk + G < n_chunks;
    // T[b] of this chunk (4 KB, L2-resident): loaded before the waits, split and staged below
    const float4 tv0 = __ldg(reinterpret_cast<const float4*>(i0.tb) + t), tv1 = __ldg(reinterpret_cast<const float4*>(i0.tb) + t + CHUNK);
    if (has_next) gather(i1, p ^ 1);                             // indices loaded an iteration ago
    if (chunk + 2 * G < n_chunks) i2 = load_idx(chunk + 2 * G);  // in flight until the next iteration
    if (has_next) cp_async_wait<1>(); else cp_async_wait<0>();
    // lo = x - trunc(x) of this thread's own 16-byte pieces (the same thread copied them: visible after its own wait_group)
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int o = q * A_LBO + (warp * 32 + 4 * it + g) * 16;
      const float4 v = *reinterpret_cast<const float4*>(s.raw[p] + o);
      float4 h4, l4;
      split_a<RN>(v.x, h4.x, l4.x), split_a<RN>(v.y, h4.y, l4.y), split_a<RN>(v.z, h4.z, l4.z), split_a<RN>(v.w, h4.w, l4.w);
      *reinterpret_cast<float4*>(s.lo + o) = l4;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // B[n][k] hi / lo, as in the one-chunk kernel
      const int f4 = t + CHUNK * k;
      const float4 v = k ? tv1 : tv0;
      const int l = f4 >> 3, m0 = (f4 & 7) * 4;
      const float vv[4] = {v.x, v.y, v.z, v.w};
      if (!TRANSPOSED) {
        float h4[4], l4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_b<RN>(vv[j], h4[j], l4[j]);
        const int o = tc::chunk_off(l, m0 / 4, D) / 4;
        *reinterpret_cast<float4*>(&s.b[0][o]) = make_float4(h4[0], h4[1], h4[2], h4[3]);
        *reinterpret_cast<float4*>(&s.b[1][o]) = make_float4(l4[0], l4[1], l4[2], l4[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a, c;
          split_b<RN>(vv[j], a, c);
          const int o = (tc::chunk_off(m0 + j, l / 4, D) + (l % 4) * 4) / 4;
          s.b[0][o] = a, s.b[1][o] = c;
        }
      }
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_thread_sync();
      const uint64_t dah = tc::make_smem_desc(tc::smem_u32(s.raw[p]), A_LBO, 128), dal = tc::make_smem_desc(tc::smem_u32(s.lo), A_LBO, 128);
      const uint64_t dbh = tc::make_smem_desc(tc::smem_u32(s.b[0]), D * 16, 128), dbl = tc::make_smem_desc(tc::smem_u32(s.b[1]), D * 16, 128);
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < D / 8; ++ks) {
          const uint64_t ao = (uint64_t)((ks * 2 * A_LBO) >> 4), bo = (uint64_t)((ks * 2 * D * 16) >> 4);
          tc::mma_tf32(tmem + p * D, dah + ao, dbh + bo, idesc, ks > 0);
          tc::mma_tf32(tmem + p * D, dah + ao, dbl + bo, idesc, true);
          tc::mma_tf32(tmem + p * D, dal + ao, dbh + bo, idesc, true);
        }
        tc::mma_commit(&s.bar[p]);
      }
      __syncwarp();
    }
    tc::mbar_wait(&s.bar[p], (uint32_t)((i >> 1) & 1));
    tc::fence_after_thread_sync();
    float v[32];
    tc::tmem_ld32(tmem + p * D + ((uint32_t)(warp * 32) << 16), v);
    float* stg = reinterpret_cast<float*>(s.raw[p]);  // the MMAs have read it; the next copy into it is issued after the sync below
#pragma unroll
    for (int c = 0; c < 32; ++c) stg[t * STG_LD + c] = i0.mult * v[c];
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float* sr = stg + (warp * 32 + 4 * it + g) * STG_LD + 4 * q;
      if (i0.gpos[it] >= 0) reinterpret_cast<float4*>(msg + (int64_t)i0.gpos[it] * D)[q] = make_float4(sr[0], sr[1], sr[2], sr[3]);
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    i0 = i1, i1 = i2;
  }
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

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

}  // namespace msg32
}  // namespace imp

using namespace imp;

extern "C" int imp_edge_messages_grouped_tc32(const imp_graph_t* g, const float* d_x, int32_t d, const float* d_table_cat,
                                              const float* d_table_an, int32_t transposed, float* d_msg, void* d_workspace,
                                              void* stream) {
  IMP_REQUIRE(g && g->n_unique >= 0 && g->bond_vocab >= 1, IMP_ERR_ARG, "imp_edge_messages_grouped_tc32: bad graph");
  IMP_REQUIRE(d == msg32::D, IMP_ERR_DIM, "imp_edge_messages_grouped_tc32: atom_dim %d not supported (32)", d);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(d_x && d_table_cat && d_table_an && d_msg && d_workspace && g->bucket_ptr && g->bucket_perm && g->col_src && g->edge_bm,
              IMP_ERR_ARG, "imp_edge_messages_grouped_tc32: null pointer");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_edge_messages_grouped_tc32: tcgen05 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * g->bond_vocab;
  int32_t* chunk_ptr = reinterpret_cast<int32_t*>(d_workspace);
  msg32::chunk_scan_kernel<<<1, 32, 0, st>>>(g->bucket_ptr, nb, chunk_ptr);
  IMP_LAUNCH_CHECK();
  const unsigned grid = (unsigned)(ceil_div(g->n_unique, msg32::CHUNK) + nb);  // upper bound; surplus CTAs exit at once
  if (transposed)
    msg32::grouped_msg_tf32x3_kernel<true><<<grid, msg32::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                          g->col_src, g->edge_bm, d_x, d_table_cat, d_table_an, d_msg);
  else
    msg32::grouped_msg_tf32x3_kernel<false><<<grid, msg32::CHUNK, 0, st>>>(g->bucket_ptr, chunk_ptr, nb, g->bond_vocab, g->bucket_perm,
                                                                           g->col_src, g->edge_bm, d_x, d_table_cat, d_table_an, d_msg);
  IMP_LAUNCH_CHECK();
  return 0;
}

/* d_plan: the index plan of imp_edge_messages_tc16_plan (imp_edge_messages_tc16_plan_bytes bytes), built once per batch. */
extern "C" int imp_edge_messages_grouped_tc32_planned(const imp_graph_t* g, const void* d_plan, const float* d_x, int32_t d,
                                                      const float* d_table_cat, const float* d_table_an, int32_t transposed,
                                                      float* d_msg, void* stream) {
  IMP_REQUIRE(g && g->n_unique >= 0, IMP_ERR_ARG, "imp_edge_messages_grouped_tc32_planned: bad graph");
  IMP_REQUIRE(d == msg32::D, IMP_ERR_DIM, "imp_edge_messages_grouped_tc32_planned: atom_dim %d not supported (32)", d);
  if (g->n_unique == 0) return 0;
  IMP_REQUIRE(d_plan && d_x && d_table_cat && d_table_an && d_msg && g->bucket_ptr && g->bucket_perm, IMP_ERR_ARG,
              "imp_edge_messages_grouped_tc32_planned: null pointer");
  IMP_REQUIRE(g->bond_vocab > 0 && 2 * g->bond_vocab <= msg32::MAX_BUCKETS, IMP_ERR_ARG,
              "imp_edge_messages_grouped_tc32_planned: bond vocabulary out of range");
  IMP_REQUIRE(imp_device_is_sm100(), IMP_ERR_UNSUPPORTED, "imp_edge_messages_grouped_tc32_planned: tcgen05 needs an sm_100 device");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * g->bond_vocab;
  const int32_t* plan = reinterpret_cast<const int32_t*>(d_plan);
  int dev = 0, sms = 148;
  IMP_CUDA(cudaGetDevice(&dev));
  IMP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid_all = (unsigned)(ceil_div(g->n_unique, msg32::CHUNK) + nb);
  const unsigned grid = grid_all < (unsigned)(3 * sms) ? grid_all : (unsigned)(3 * sms);
  const size_t smem = sizeof(msg32::PSmem) + 128;
#define IMP_MSG32_LAUNCH(TR, RN)                                                                                                            \
  do {                                                                                                                                     \
    IMP_CUDA(cudaFuncSetAttribute(msg32::grouped_msg_tf32x3_planned_kernel<TR, RN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    msg32::grouped_msg_tf32x3_planned_kernel<TR, RN><<<grid, msg32::CHUNK, smem, st>>>(g->bucket_ptr, plan, nb, g->bond_vocab, g->bucket_perm, \
                                                                                       plan + 1024, plan + 1024 + g->n_unique, d_x,       \
                                                                                       d_table_cat, d_table_an, d_msg);                  \
  } while (0)
  const bool tr = (transposed & 1) != 0, rn = (transposed & 2) != 0;  // bit 1: rounded operand splits (the fp32 inference route)
  if (tr && rn) IMP_MSG32_LAUNCH(true, true);
  else if (tr) IMP_MSG32_LAUNCH(true, false);
  else if (rn) IMP_MSG32_LAUNCH(false, true);
  else IMP_MSG32_LAUNCH(false, false);
#undef IMP_MSG32_LAUNCH
  IMP_LAUNCH_CHECK();
  return 0;
}
