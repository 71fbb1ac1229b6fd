// imp_fused_plan: cuts a packed graph batch into the 128-row tile records of fused_plan.cuh.
//
// Replaces, for the fused forward, what src/dataset.py / build_inputs (train_viscosity.py:291-314) do for the reference:
// decide which ions share a device batch row block.  Integer-only, deterministic per tile (the ORDER of a window's tiles in
// the plan depends on an atomic counter; the contents of every tile, and therefore every result, do not).
//
// One warp per window of FP_WIN consecutive molecules of one tower:
//   1. lanes load the molecules' atom and entry counts;
//   2. lane 0 packs them best-fit (largest molecule that still fits the rows and entries left, size buckets + a bit mask);
//   3. the warp emits the window's tiles: per tile a scan over the chosen molecules, four 32-row passes (atom id, in-degree,
//      entry offset, counting sort by in-degree with ballots), a coalesced copy of every molecule's contiguous CSR range
//      translated to tile rows, and a coalesced 2 KiB store of the record.
#include <cuda_fp16.h>

#include "common.cuh"
#include "fused_plan.cuh"

namespace imp {

constexpr int PL_WARPS = 4;

struct PlanArgs {
  const int* mol_ptr;
  const int* atom_id;
  const int* row_ptr;
  const int* col_src;
  const int* edge_bm;
  const int* mol_eptr;
  const unsigned short* atom_w;
  const unsigned int* edge_w;
  const unsigned short* edge_h;  // 16-bit entry words (imp_fused_plan_compact16): src | bond << 7 | (multiplicity - 1) << 15
  unsigned char* plan;
  int n_pairs, atom_vocab, bond_vocab;
};

struct alignas(16) PlanWarpSmem {
  FusedTile tile;
  int aptr[FP_WIN];  // first atom of every molecule of the window
  int eptr[FP_WIN];  // first CSR entry
  unsigned short ments[FP_WIN];
  unsigned short next[FP_WIN];
  unsigned short order[FP_WIN];
  unsigned short tstart[FP_WIN + 2];
  unsigned short head[FP_ROWS + 2];
  unsigned char msize[FP_WIN];
};

constexpr unsigned short PL_NIL = 0xffff;

__device__ __forceinline__ unsigned int half_bits_of_int(int v) { return (unsigned int)__half_as_ushort(__float2half_rn((float)v)); }
__device__ __forceinline__ int field8(unsigned long long v, int k) { return (int)((v >> (8 * k)) & 0xffull); }

template <int FEED>  // 0: int32 CSR arrays, 1: compact feed (32-bit entry words), 2: compact feed with 16-bit entry words
__global__ void __launch_bounds__(PL_WARPS * 32) fused_plan_kernel(const PlanArgs a) {
  constexpr bool COMPACT = FEED != 0;
  __shared__ PlanWarpSmem sm[PL_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PlanWarpSmem& ws = sm[warp];
  FusedPlanHeader* hdr = reinterpret_cast<FusedPlanHeader*>(a.plan);
  const int P = a.n_pairs;
  const int nwin = (P + FP_WIN - 1) / FP_WIN;
  const int w = blockIdx.x * PL_WARPS + warp;
  if (w >= 2 * nwin) return;
  const int tower = w >= nwin;
  const int m0 = (tower ? w - nwin : w) * FP_WIN;  // first molecule of the window inside its tower
  const int nw = min(FP_WIN, P - m0);
  const int gbase = tower * P + m0;  // global (tower-major) molecule index of the window's first molecule
  int bad = 0;

  // ---- 1. sizes
  for (int i = lane; i < FP_WIN; i += 32) {
    int sz = 0, me = 0, p0 = 0, e0 = 0;
    if (i < nw) {
      p0 = __ldg(a.mol_ptr + gbase + i);
      const int p1 = __ldg(a.mol_ptr + gbase + i + 1);
      sz = p1 - p0;
      int e1;
      if (COMPACT) e0 = __ldg(a.mol_eptr + gbase + i), e1 = __ldg(a.mol_eptr + gbase + i + 1);
      else e0 = __ldg(a.row_ptr + p0), e1 = __ldg(a.row_ptr + p1);
      me = e1 - e0;
      if (sz < 0 || sz > FP_ROWS || me < 0 || me > FP_ECAP) bad = 1, sz = min(max(sz, 0), FP_ROWS), me = min(max(me, 0), FP_ECAP);
    }
    ws.msize[i] = (unsigned char)sz;
    ws.ments[i] = (unsigned short)me;
    ws.aptr[i] = p0, ws.eptr[i] = e0;
  }
  for (int i = lane; i < FP_ROWS + 2; i += 32) ws.head[i] = PL_NIL;
  __syncwarp();

  // ---- 2. best-fit packing (lane 0)
  int nt = 0;
  if (lane == 0) {
    unsigned long long mlo = 0ull, mhi = 0ull;  // bit s-1 of mlo: a molecule of s atoms (1..64) is left; mhi: 65..128
    bool zero = false;
    for (int i = nw - 1; i >= 0; --i) {  // chains pop in ascending molecule order
      const int s = ws.msize[i];
      ws.next[i] = ws.head[s];
      ws.head[s] = (unsigned short)i;
      if (s == 0) zero = true;
      else if (s <= 64) mlo |= 1ull << (s - 1);
      else mhi |= 1ull << (s - 65);
    }
    int pos = 0;
    while (pos < nw) {
      ws.tstart[nt] = (unsigned short)pos;
      int gap = FP_ROWS, egap = FP_ECAP, nmol = 0;
      while (nmol < FP_MAXMOL) {
        int s = -1;  // largest size <= gap that is left
        if (gap > 64) {
          const unsigned long long m = gap >= 128 ? mhi : (mhi & ((1ull << (gap - 64)) - 1ull));
          if (m) s = 128 - __clzll((long long)m);
        }
        if (s < 0) {
          const int g2 = min(gap, 64);
          const unsigned long long m = g2 >= 64 ? mlo : (mlo & ((1ull << g2) - 1ull));
          if (m) s = 64 - __clzll((long long)m);
        }
        if (s < 0 && zero) s = 0;
        if (s < 0) break;
        const int i = ws.head[s];
        if ((int)ws.ments[i] > egap) break;  // (a lone molecule always fits: ments <= FP_ECAP was enforced above)
        ws.head[s] = ws.next[i];
        if (ws.head[s] == PL_NIL) {
          if (s == 0) zero = false;
          else if (s <= 64) mlo &= ~(1ull << (s - 1));
          else mhi &= ~(1ull << (s - 65));
        }
        ws.order[pos++] = (unsigned short)i;
        gap -= s, egap -= ws.ments[i], ++nmol;
      }
      ++nt;
    }
    ws.tstart[nt] = (unsigned short)pos;
  }
  nt = __shfl_sync(0xffffffffu, nt, 0);
  int tbase = 0;
  if (lane == 0) tbase = atomicAdd(&hdr->n_tiles[tower], nt);
  tbase = __shfl_sync(0xffffffffu, tbase, 0);
  const int cap = hdr->cap[tower];
  if (tbase + nt > cap) {
    if (lane == 0) atomicMax(&hdr->status, 2);
    return;
  }
  FusedTile* out = reinterpret_cast<FusedTile*>(a.plan + FP_HEADER_BYTES) + (size_t)(tower ? hdr->cap[0] : 0) + tbase;
  __syncwarp();

  // ---- 3. tiles
  for (int k = 0; k < nt; ++k) {
    const int o0 = ws.tstart[k], nmol = ws.tstart[k + 1] - o0;
    int sz = 0, me = 0, aptr = 0, eptr = 0;
    if (lane < nmol) {
      const int i = ws.order[o0 + lane];
      sz = ws.msize[i], me = ws.ments[i], aptr = ws.aptr[i], eptr = ws.eptr[i];
      ws.tile.molid[lane] = gbase + i;
    }
    int off = sz, eoff = me;  // inclusive scans over the tile's molecules
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, off, o), u = __shfl_up_sync(0xffffffffu, eoff, o);
      if (lane >= o) off += v, eoff += u;
    }
    const int rows = __shfl_sync(0xffffffffu, off, 31), n_ent = __shfl_sync(0xffffffffu, eoff, 31);
    const int aend = lane < nmol ? off : 0x7fffffff;  // first natural row after molecule `lane`
    off -= sz, eoff -= me;
    if (lane < nmol) ws.tile.mol_lo[lane] = (uint8_t)off;
    if (lane == 0) {
      ws.tile.mol_lo[nmol] = (uint8_t)rows;
      ws.tile.nm = (uint8_t)nmol, ws.tile.rows = (uint8_t)rows, ws.tile.n_ent = (uint16_t)n_ent;
    }
    // rows: four passes of 32 natural rows, in three sweeps so that every global load of the tile's rows is in flight before
    // the first is used (the kernel is bound by load latency: ncu long scoreboard).
    int jm[4], at[4];  // molecule (lane index in this tile) and global atom of natural row 32 p + lane
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int rho = 32 * p + lane;
      int j = 0;  // number of molecules that end at or before the row
      for (int q = 0; q < nmol; ++q) j += (__shfl_sync(0xffffffffu, aend, q) <= rho) ? 1 : 0;
      jm[p] = min(j, 31);
      at[p] = __shfl_sync(0xffffffffu, aptr, jm[p]) + (rho - __shfl_sync(0xffffffffu, off, jm[p]));
    }
    int aidv[4], degv[4], g0[4];  // atom id, in-degree, first global CSR entry of the row
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      aidv[p] = 0, degv[p] = 0, g0[p] = 0;
      if (32 * p + lane < rows) {
        if (COMPACT) {
          const int aw = (int)__ldg(a.atom_w + at[p]);
          aidv[p] = aw & 0xff, degv[p] = aw >> 8;
        } else {
          aidv[p] = __ldg(a.atom_id + at[p]);
          g0[p] = __ldg(a.row_ptr + at[p]);
          degv[p] = __ldg(a.row_ptr + at[p] + 1);
        }
      }
    }
    // In-degree classes (min(deg, 7)) are counted in eight 8-bit fields of one 64-bit word: an inclusive warp scan of
    // "1 << 8 key" gives every row its rank inside its class (stable, natural order).
    int word[4], key[4], e0v[4];
    unsigned long long incl_cls[4];
    unsigned long long run_cls = 0ull;  // class counts of the passes before this one
    int carry = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int rho = 32 * p + lane;
      int deg = degv[p], aid = aidv[p];
      if (rho < rows) {
        if (!COMPACT) deg -= g0[p];
        if (deg < 0 || deg > 31) bad = 1, deg = min(max(deg, 0), 31);
        aid = min(max(aid, 0), min(a.atom_vocab - 1, 1023));
      }
      key[p] = min(deg, 7);
      int incl = deg;
      unsigned long long cls = 1ull << (8 * key[p]);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        const unsigned long long c = __shfl_up_sync(0xffffffffu, cls, o);
        if (lane >= o) incl += v, cls += c;
      }
      const int e0 = min(carry + incl - deg, FP_ECAP);  // (clamps only bite on refused batches: no read leaves the record)
      carry += __shfl_sync(0xffffffffu, incl, 31);
      deg = min(deg, FP_ECAP - e0);
      incl_cls[p] = cls + run_cls;
      run_cls += __shfl_sync(0xffffffffu, cls, 31);
      word[p] = rho | (deg << 7) | (e0 << 12) | (aid << 22);
      degv[p] = deg, e0v[p] = e0;
      if (COMPACT)  // entries of a molecule are contiguous: first entry of the molecule + entries of its earlier rows
        g0[p] = __shfl_sync(0xffffffffu, eptr, jm[p]) + (e0 - __shfl_sync(0xffffffffu, eoff, jm[p]));
    }
    // exclusive prefix over the classes of the tile totals (<= 128 per field: no carry between fields)
    const unsigned long long cls_base = run_cls * 0x0101010101010100ull;
#pragma unroll
    for (int p = 0; p < 4; ++p) ws.tile.slot[field8(cls_base, key[p]) + field8(incl_cls[p], key[p]) - 1] = (uint32_t)word[p];
    // entries, row-parallel: lane (p, lane) copies its row's entries, translated to tile rows; the first four entries of
    // all four passes are loaded before any is used
    {
      int rowbase[4], msz[4];  // natural row of atom 0 of the row's molecule, atoms of that molecule
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        rowbase[p] = __shfl_sync(0xffffffffu, off, jm[p]);
        msz[p] = __shfl_sync(0xffffffffu, sz, jm[p]);
      }
      auto put = [&](int p, int i, int csrc, int cbm, unsigned int cw) {
        int src, bond, mult;
        if (FEED == 2) src = (int)(cw & 0x7fu), bond = (int)((cw >> 7) & 0xffu), mult = (int)(cw >> 15) + 1;
        else if (COMPACT) src = (int)(cw & 0xffu), bond = (int)((cw >> 8) & 0xffu), mult = (int)((cw >> 16) & 0xffu);
        else src = csrc - (at[p] - (32 * p + lane - rowbase[p])), bond = cbm & 0xffff, mult = cbm >> 16;
        if (src < 0 || src >= msz[p]) bad = 1, src = min(max(src, 0), max(msz[p] - 1, 0));
        bond = min(bond, min(a.bond_vocab - 1, 255));
        ws.tile.ent[e0v[p] + i] = (uint32_t)(src + rowbase[p]) | ((uint32_t)bond << 8) | (half_bits_of_int(mult) << 16);
      };
      int csrc[4][4], cbm[4][4];
      unsigned int cw[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          csrc[p][i] = 0, cbm[p][i] = 0, cw[p][i] = 0u;
          if (i < degv[p]) {
            if (FEED == 2) cw[p][i] = (unsigned int)__ldg(a.edge_h + g0[p] + i);
            else if (COMPACT) cw[p][i] = __ldg(a.edge_w + g0[p] + i);
            else csrc[p][i] = __ldg(a.col_src + g0[p] + i), cbm[p][i] = __ldg(a.edge_bm + g0[p] + i);
          }
        }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < degv[p]) put(p, i, csrc[p][i], cbm[p][i], cw[p][i]);
        for (int i = 4; i < degv[p]; ++i) {  // rows with more than four entries
          if (FEED == 2) put(p, i, 0, 0, (unsigned int)__ldg(a.edge_h + g0[p] + i));
          else if (COMPACT) put(p, i, 0, 0, __ldg(a.edge_w + g0[p] + i));
          else put(p, i, __ldg(a.col_src + g0[p] + i), __ldg(a.edge_bm + g0[p] + i), 0u);
        }
      }
    }
    __syncwarp();
    {
      const uint4* s4 = reinterpret_cast<const uint4*>(&ws.tile);
      uint4* d4 = reinterpret_cast<uint4*>(out + k);
#pragma unroll
      for (int q = 0; q < (int)sizeof(FusedTile) / 16 / 32; ++q) d4[q * 32 + lane] = s4[q * 32 + lane];
    }
    __syncwarp();
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicMax(&hdr->status, 1);
}

__global__ void fused_plan_init_kernel(FusedPlanHeader* hdr, int cap0, int cap1) {
  hdr->n_tiles[0] = 0, hdr->n_tiles[1] = 0, hdr->cap[0] = cap0, hdr->cap[1] = cap1, hdr->status = 0;
}

// Upper bound on the tiles of one tower with `atoms` atoms, `ents` entries and `mols` molecules of at most `max_mol` atoms:
// every tile of a window but its last is closed because the next molecule did not fit -- by rows (> 128 - max_mol used),
// by entries (> FP_ECAP - max entries of a molecule used; bounded through the half-capacity argument) or by the molecule
// count -- plus one partial tile per window.
static int64_t plan_tile_bound(int64_t atoms, int64_t ents, int64_t mols, int max_mol) {
  const int64_t by_rows = atoms / (FP_ROWS - (max_mol < 1 ? 1 : max_mol > 127 ? 127 : max_mol) + 1) + 1;
  const int64_t by_ents = ents / (FP_ECAP / 2) + 1;
  const int64_t by_mols = mols / FP_MAXMOL + 1;
  const int64_t windows = (mols + FP_WIN - 1) / FP_WIN;
  int64_t t = by_rows + by_ents + by_mols + windows;
  return t < mols + 1 ? t : mols + 1;
}

}  // namespace imp

using namespace imp;

extern "C" int64_t imp_fused_plan_bytes(int32_t n_pairs, int32_t n_atoms, int32_t n_unique, int32_t max_mol_atoms) {
  if (n_pairs < 0 || n_atoms < 0 || n_unique < 0) return IMP_ERR_ARG;
  // either tower may hold all atoms / entries of the batch; sized for the worst split
  const int64_t cap = plan_tile_bound(n_atoms, n_unique, n_pairs, max_mol_atoms);
  return FP_HEADER_BYTES + 2 * cap * (int64_t)sizeof(FusedTile);
}

static int fused_plan_impl(const imp_graph_t* g, const imp_compact_graph_t* cg, bool narrow, int32_t atom_vocab, int32_t max_mol_atoms,
                           void* d_plan, int64_t plan_bytes, void* stream) {
  IMP_REQUIRE((g != nullptr) != (cg != nullptr), IMP_ERR_ARG, "imp_fused_plan: pass exactly one of the two graph forms");
  PlanArgs a{};
  int n_atoms, n_unique;
  if (cg) {
    IMP_REQUIRE(cg->mol_ptr && cg->mol_eptr && (cg->n_atoms == 0 || cg->atom_w) && (cg->n_unique == 0 || cg->edge_w), IMP_ERR_ARG,
                "imp_fused_plan: null index arrays");
    a.mol_ptr = cg->mol_ptr, a.mol_eptr = cg->mol_eptr, a.atom_w = cg->atom_w, a.edge_w = cg->edge_w;
    a.edge_h = reinterpret_cast<const unsigned short*>(cg->edge_w);
    a.n_pairs = cg->n_pairs, a.bond_vocab = cg->bond_vocab, n_atoms = cg->n_atoms, n_unique = cg->n_unique;
  } else {
    IMP_REQUIRE(g->mol_ptr && g->row_ptr && (g->n_atoms == 0 || g->atom_id) && (g->n_unique == 0 || (g->col_src && g->edge_bm)),
                IMP_ERR_ARG, "imp_fused_plan: null index arrays");
    a.mol_ptr = g->mol_ptr, a.atom_id = g->atom_id, a.row_ptr = g->row_ptr, a.col_src = g->col_src, a.edge_bm = g->edge_bm;
    a.n_pairs = g->n_pairs, a.bond_vocab = g->bond_vocab, n_atoms = g->n_atoms, n_unique = g->n_unique;
  }
  IMP_REQUIRE(a.n_pairs >= 0 && n_atoms >= 0 && n_unique >= 0 && atom_vocab >= 1 && a.bond_vocab >= 1, IMP_ERR_ARG, "imp_fused_plan: bad sizes");
  IMP_REQUIRE(max_mol_atoms <= FP_ROWS, IMP_ERR_DIM, "imp_fused_plan: a molecule has %d atoms, a tile holds %d", max_mol_atoms, FP_ROWS);
  IMP_REQUIRE(d_plan && plan_bytes >= FP_HEADER_BYTES, IMP_ERR_ARG, "imp_fused_plan: no plan buffer");
  const int64_t cap = (plan_bytes - FP_HEADER_BYTES) / (2 * (int64_t)sizeof(FusedTile));
  IMP_REQUIRE(cap >= 1 || a.n_pairs == 0, IMP_ERR_CAPACITY, "imp_fused_plan: plan buffer too small");
  a.plan = (unsigned char*)d_plan, a.atom_vocab = atom_vocab;
  cudaStream_t st = (cudaStream_t)stream;
  fused_plan_init_kernel<<<1, 1, 0, st>>>(reinterpret_cast<FusedPlanHeader*>(d_plan), (int)(cap > INT32_MAX ? INT32_MAX : cap),
                                         (int)(cap > INT32_MAX ? INT32_MAX : cap));
  IMP_LAUNCH_CHECK();
  if (a.n_pairs == 0) return 0;
  const int nwin = (a.n_pairs + FP_WIN - 1) / FP_WIN;
  const int grid = (2 * nwin + PL_WARPS - 1) / PL_WARPS;
  if (cg && narrow) fused_plan_kernel<2><<<grid, PL_WARPS * 32, 0, st>>>(a);
  else if (cg) fused_plan_kernel<1><<<grid, PL_WARPS * 32, 0, st>>>(a);
  else fused_plan_kernel<0><<<grid, PL_WARPS * 32, 0, st>>>(a);
  IMP_LAUNCH_CHECK();
  return 0;
}

extern "C" int imp_fused_plan(const imp_graph_t* g, const imp_compact_graph_t* cg, int32_t atom_vocab, int32_t max_mol_atoms,
                              void* d_plan, int64_t plan_bytes, void* stream) {
  return fused_plan_impl(g, cg, false, atom_vocab, max_mol_atoms, d_plan, plan_bytes, stream);
}

extern "C" int imp_fused_plan_compact16(const imp_compact_graph_t* cg, int32_t atom_vocab, int32_t max_mol_atoms, void* d_plan,
                                        int64_t plan_bytes, void* stream) {
  IMP_REQUIRE(cg, IMP_ERR_ARG, "imp_fused_plan_compact16: graph is null");
  return fused_plan_impl(nullptr, cg, true, atom_vocab, max_mol_atoms, d_plan, plan_bytes, stream);
}
