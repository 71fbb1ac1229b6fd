// Device-side packed-CSR builder ("GPU-side packer + input feed").
//
// Same contract as imp_pack_host (pack_host.cpp; bit-exact specification oracle/ref_pack.py): replaces
// pad_sequences_1d (train_viscosity.py:52-59), preprocess_edges_and_bonds (:76-110), the +1 id shifts (:255-262) and the
// masks of BondMatrixMessage / Reduce (models/layers.py:114-115, 74-76).  Input = the flat ragged ion arrays already on
// the device; output = mol_ptr, atom_id, row_ptr, col_src, edge_bm (what the fused forward reads; the bond-bucket
// permutation of the staged / training kernels is not produced here) and, optionally, the compact feed.
//
// One warp per molecule, two passes with no temporary storage:
//   pass A  count the unique live entries of the molecule (an entry is "first" if no earlier identical one exists)
//           and validate indices; exclusive scan of the counts over molecules (three small kernels)
//   pass B  rank-sort the molecule's live entries by (dst, bond, src) in shared memory (O(n^2) compares per molecule,
//           n ~ 100: ~10^4 per warp), merge duplicates into a multiplicity, write the CSR rows at the scanned offset.
// Molecules with more than PD_CAP entries after doubling set status = IMP_ERR_CAPACITY (use imp_pack_host).
#include "common.cuh"

namespace imp {

constexpr int PD_CAP = 512;     // doubled entries per molecule held in shared memory
constexpr int PD_WARPS = 8;     // molecules per CTA
constexpr int PD_MAX_ATOMS = 1024;

struct PackDevArgs {
  imp_ions_t cat, an;  // device pointers inside
  int n_pairs, n_cat_atoms, bond_vocab, max_edges, flags;
  int* status;         // 0 = ok, else an IMP_ERR_* code (first error wins is not guaranteed; any error is fatal)
  int* uniq;           // [2P+1] pass A: counts; after the scan: offsets
  // outputs
  int* mol_ptr;
  int* atom_id;
  int* row_ptr;
  int* col_src;
  int* edge_bm;
  unsigned long long* n_edges;  // sum of multiplicities
  int edge_capacity;
  // optional compact feed
  int* mol_eptr;
  unsigned short* atom_w;
  unsigned int* edge_w;
};

// Loads the doubled / truncated entry list of molecule m into keys[] (dst << 40 | bond << 24 | src; live entries only,
// dead ones are skipped).  Returns the number of live entries, or -1 after setting *status.
__device__ int pd_load_keys(const PackDevArgs& a, int m, unsigned long long* keys, int lane, int* n_atoms_out, int* atom_base_in) {
  const bool is_cat = m < a.n_pairs;
  const imp_ions_t& ions = is_cat ? a.cat : a.an;
  const int i = is_cat ? m : m - a.n_pairs;
  const int n = ions.atom_ptr[i + 1] - ions.atom_ptr[i];
  *n_atoms_out = n;
  *atom_base_in = ions.atom_ptr[i];
  const int e0 = ions.edge_ptr[i], ne = ions.edge_ptr[i + 1] - e0;
  const bool dbl = a.flags & IMP_PACK_DOUBLE_EDGES;
  const int shift = (a.flags & IMP_PACK_SHIFT_IDS) ? 1 : 0;
  long long total = dbl ? 2LL * ne : ne;
  if (a.max_edges >= 0 && total > 2LL * a.max_edges) total = 2LL * a.max_edges;
  if (total > PD_CAP) {
    if (lane == 0) atomicExch(a.status, IMP_ERR_CAPACITY);
    return -1;
  }
  // produced entry p comes from input entry p / 2 (doubled) with direction p % 2; compact the live ones in order
  int n_live = 0;
  bool bad = false;
  for (int p0 = 0; p0 < (int)total; p0 += 32) {
    const int p = p0 + lane;
    bool live = false;
    unsigned long long key = 0;
    if (p < (int)total) {
      const int e = e0 + (dbl ? p / 2 : p);
      const bool rev = dbl && (p & 1);
      const int s = ions.edge_src[e], t = ions.edge_dst[e], b = ions.bond_ids[e] + shift;
      const int src = rev ? t : s, dst = rev ? s : t;
      if (src > 0 && dst > 0) {
        if (src >= n || dst >= n || b <= 0 || b >= a.bond_vocab) bad = true;
        live = true;
        key = ((unsigned long long)dst << 40) | ((unsigned long long)b << 24) | (unsigned long long)src;
      } else if (src < 0 || dst < 0) {
        bad = true;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (live) keys[n_live + __popc(mask & ((1u << lane) - 1u))] = key;
    n_live += __popc(mask);
  }
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) atomicExch(a.status, IMP_ERR_INDEX);
    return -1;
  }
  __syncwarp();
  return n_live;
}

__global__ void __launch_bounds__(PD_WARPS * 32) pack_count_kernel(PackDevArgs a) {
  __shared__ unsigned long long skeys[PD_WARPS][PD_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * PD_WARPS + warp;
  if (m >= 2 * a.n_pairs) return;
  unsigned long long* keys = skeys[warp];
  int n_atoms, abase;
  const int n = pd_load_keys(a, m, keys, lane, &n_atoms, &abase);
  int uniq = 0;
  if (n > 0) {
    for (int i = lane; i < n; i += 32) {
      const unsigned long long k = keys[i];
      bool first = true;
      for (int j = 0; j < i; ++j)
        if (keys[j] == k) {
          first = false;
          break;
        }
      uniq += first;
    }
    for (int s = 16; s > 0; s >>= 1) uniq += __shfl_xor_sync(0xffffffffu, uniq, s);
  }
  if (lane == 0) {
    a.uniq[m] = n > 0 ? uniq : 0;
    if (n_atoms < 0 || n_atoms > PD_MAX_ATOMS) atomicExch(a.status, IMP_ERR_CAPACITY);
  }
}

// ---- exclusive scan of int32 (n up to ~2^31 / 1024 blocks): block sums, scan of block sums, local scan + offset
constexpr int SC_BLOCK = 1024;
__global__ void __launch_bounds__(256) scan_block_sums_kernel(const int* __restrict__ x, int n, long long* __restrict__ sums) {
  __shared__ long long red[256];
  const int base = blockIdx.x * SC_BLOCK;
  long long s = 0;
  for (int i = threadIdx.x; i < SC_BLOCK; i += 256)
    if (base + i < n) s += x[base + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = red[0];
}
__global__ void scan_sums_kernel(long long* sums, int n_blocks) {  // one thread: n_blocks is small (<= a few thousand)
  long long run = 0;
  for (int i = 0; i < n_blocks; ++i) {
    const long long v = sums[i];
    sums[i] = run;
    run += v;
  }
  sums[n_blocks] = run;
}
__global__ void __launch_bounds__(SC_BLOCK) scan_apply_kernel(int* __restrict__ x, int n, const long long* __restrict__ sums,
                                                              int* __restrict__ status, int capacity) {
  __shared__ int wsum[32];
  const int i = blockIdx.x * SC_BLOCK + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = i < n ? x[i] : 0;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const long long off = sums[blockIdx.x] + (warp > 0 ? wsum[warp - 1] : 0) + incl - v;
  if (i < n) x[i] = (int)off;
  if (i == n - 1 && (off + v > capacity || off + v > 0x7fffffffLL)) atomicExch(status, IMP_ERR_CAPACITY);
  if (i == n - 1) x[n] = (int)(off + v);  // total in the extra slot
}

constexpr int PD_WRITE_SMEM = PD_WARPS * (2 * PD_CAP * 8 + PD_MAX_ATOMS * 4);  // 96 KiB, dynamic

__global__ void __launch_bounds__(PD_WARPS * 32) pack_write_kernel(PackDevArgs a) {
  extern __shared__ __align__(16) unsigned char pd_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * PD_WARPS + warp;
  if (m >= 2 * a.n_pairs) return;
  if (*a.status != 0) return;  // pass A or the scan failed: outputs stay undefined, the host reports the error
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(pd_smem) + (size_t)warp * PD_CAP;
  unsigned long long* sorted = reinterpret_cast<unsigned long long*>(pd_smem) + (size_t)(PD_WARPS + warp) * PD_CAP;
  int* deg = reinterpret_cast<int*>(pd_smem + (size_t)2 * PD_WARPS * PD_CAP * 8) + (size_t)warp * PD_MAX_ATOMS;
  int n_atoms, abase;
  const int n = pd_load_keys(a, m, keys, lane, &n_atoms, &abase);
  const bool is_cat = m < a.n_pairs;
  const imp_ions_t& ions = is_cat ? a.cat : a.an;
  const int i_ion = is_cat ? m : m - a.n_pairs;
  const int mol_base = (is_cat ? 0 : a.n_cat_atoms) + ions.atom_ptr[i_ion];  // global index of the molecule's atom 0
  const int shift = (a.flags & IMP_PACK_SHIFT_IDS) ? 1 : 0;
  if (lane == 0) {
    a.mol_ptr[m] = mol_base;
    if (m == 2 * a.n_pairs - 1) a.mol_ptr[m + 1] = mol_base + n_atoms;
  }
  for (int k = lane; k < n_atoms; k += 32) deg[k] = 0;
  __syncwarp();
  const int eoff = a.uniq[m];
  if (lane == 0 && a.mol_eptr) {
    a.mol_eptr[m] = eoff;
    if (m == 2 * a.n_pairs - 1) a.mol_eptr[m + 1] = a.uniq[m + 1];
  }
  unsigned long long edges = 0;
  if (n > 0) {
    // rank sort (stable): position = #keys smaller + #equal keys that come earlier
    for (int i = lane; i < n; i += 32) {
      const unsigned long long k = keys[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const unsigned long long kj = keys[j];
        rank += (kj < k) || (kj == k && j < i);
      }
      sorted[rank] = k;
    }
    __syncwarp();
    // unique runs -> CSR entries
    int written = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
      const int p = p0 + lane;
      const bool first = p < n && (p == 0 || sorted[p - 1] != sorted[p]);
      const unsigned mask = __ballot_sync(0xffffffffu, first);
      if (first) {
        const unsigned long long k = sorted[p];
        int mult = 1;
        while (p + mult < n && sorted[p + mult] == k) ++mult;
        const int dst = (int)(k >> 40), bond = (int)((k >> 24) & 0xffff), src = (int)(k & 0xffffff);
        const int pos = eoff + written + __popc(mask & ((1u << lane) - 1u));
        if (mult >= (1 << 15)) atomicExch(a.status, IMP_ERR_CAPACITY);
        if (pos < a.edge_capacity) {
          a.col_src[pos] = mol_base + src;
          a.edge_bm[pos] = bond | (mult << 16);
          if (a.edge_w) a.edge_w[pos] = (unsigned)src | ((unsigned)bond << 8) | ((unsigned)mult << 16);
        }
        atomicAdd(&deg[dst], 1);
        edges += mult;
      }
      written += __popc(mask);
    }
    __syncwarp();
  }
  for (int s = 16; s > 0; s >>= 1) edges += __shfl_xor_sync(0xffffffffu, edges, s);
  if (lane == 0 && edges) atomicAdd(a.n_edges, edges);
  // atoms: ids, row pointers (exclusive scan of the in-degrees inside the molecule)
  int run = eoff;
  for (int k0 = 0; k0 < n_atoms; k0 += 32) {
    const int k = k0 + lane;
    const int d = k < n_atoms ? deg[k] : 0;
    int incl = d;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (k < n_atoms) {
      const int id = ions.atom_ids[abase + k] + shift;
      a.atom_id[mol_base + k] = id;
      a.row_ptr[mol_base + k] = run + incl - d;
      if (a.atom_w) a.atom_w[mol_base + k] = (unsigned short)((id & 0xff) | (d << 8));
      if (a.atom_w && (id < 0 || id > 255 || d > 255)) atomicExch(a.status, IMP_ERR_DIM);
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0 && m == 2 * a.n_pairs - 1) a.row_ptr[mol_base + n_atoms] = run;
}

}  // namespace imp

using namespace imp;

extern "C" int64_t imp_pack_device_workspace_bytes(int32_t n_pairs) {
  const int64_t n = 2 * (int64_t)n_pairs + 1;
  return ((n + 1) * 4 + 15) / 16 * 16 + (ceil_div(n, SC_BLOCK) + 2) * 8 + 64;
}

// d_counts[4] (device, written): n_unique, status, n_edges (low, high 32 bits)
extern "C" int imp_pack_device(const imp_ions_t* cation, const imp_ions_t* anion, int32_t n_cat_atoms, int32_t n_atoms,
                               int32_t bond_vocab, int32_t max_edges, int32_t flags, int32_t edge_capacity, imp_graph_t* out,
                               int32_t* d_mol_eptr, uint16_t* d_atom_w, uint32_t* d_edge_w, int32_t* d_counts,
                               void* d_workspace, void* stream) {
  IMP_REQUIRE(cation && anion && out && d_counts && d_workspace, IMP_ERR_ARG, "imp_pack_device: null argument");
  IMP_REQUIRE(cation->n_ions == anion->n_ions && cation->n_ions >= 0 && bond_vocab > 0 && bond_vocab <= 0xFFFF, IMP_ERR_ARG,
              "imp_pack_device: n_ions mismatch or bad bond_vocab");
  IMP_REQUIRE(n_cat_atoms >= 0 && n_atoms >= n_cat_atoms && edge_capacity >= 0, IMP_ERR_ARG, "imp_pack_device: bad sizes");
  const int P = cation->n_ions;
  out->n_pairs = P, out->n_atoms = n_atoms, out->n_cat_atoms = n_cat_atoms, out->bond_vocab = bond_vocab;
  cudaStream_t st = (cudaStream_t)stream;
  IMP_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int32_t), st));
  if (P == 0) return 0;
  IMP_REQUIRE(out->mol_ptr && out->atom_id && out->row_ptr && (edge_capacity == 0 || (out->col_src && out->edge_bm)), IMP_ERR_ARG,
              "imp_pack_device: output arrays missing");
  PackDevArgs a;
  a.cat = *cation, a.an = *anion;
  a.n_pairs = P, a.n_cat_atoms = n_cat_atoms, a.bond_vocab = bond_vocab, a.max_edges = max_edges, a.flags = flags;
  a.status = d_counts + 1;
  a.n_edges = reinterpret_cast<unsigned long long*>(d_counts + 2);
  a.uniq = reinterpret_cast<int*>(d_workspace);
  const int n = 2 * P;
  long long* sums = reinterpret_cast<long long*>(reinterpret_cast<char*>(d_workspace) + ((int64_t)(n + 2) * 4 + 15) / 16 * 16);
  a.mol_ptr = out->mol_ptr, a.atom_id = out->atom_id, a.row_ptr = out->row_ptr, a.col_src = out->col_src, a.edge_bm = out->edge_bm;
  a.edge_capacity = edge_capacity;
  a.mol_eptr = d_mol_eptr, a.atom_w = d_atom_w, a.edge_w = d_edge_w;
  const int blocks = (int)ceil_div(n, PD_WARPS);
  pack_count_kernel<<<blocks, PD_WARPS * 32, 0, st>>>(a);
  IMP_LAUNCH_CHECK();
  const int nb = (int)ceil_div(n, SC_BLOCK);
  scan_block_sums_kernel<<<nb, 256, 0, st>>>(a.uniq, n, sums);
  IMP_LAUNCH_CHECK();
  scan_sums_kernel<<<1, 1, 0, st>>>(sums, nb);
  IMP_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, SC_BLOCK, 0, st>>>(a.uniq, n, sums, a.status, edge_capacity);
  IMP_LAUNCH_CHECK();
  IMP_CUDA(cudaFuncSetAttribute(pack_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_WRITE_SMEM));
  pack_write_kernel<<<blocks, PD_WARPS * 32, PD_WRITE_SMEM, st>>>(a);
  IMP_LAUNCH_CHECK();
  IMP_CUDA(cudaMemcpyAsync(d_counts, a.uniq + n, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));  // n_unique
  return 0;
}
