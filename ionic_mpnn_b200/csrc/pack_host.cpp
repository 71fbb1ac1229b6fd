// Host-side packed-CSR builder and the benchmark-sized synthetic ion generator.
//
// imp_pack_host replaces the reference's padding pipeline -- pad_sequences_1d (train_viscosity.py:52-59),
// preprocess_edges_and_bonds (:76-110), the +1 shifts (:255-262) -- and folds in the masks that
// BondMatrixMessage (models/layers.py:114-115) and Reduce (models/layers.py:74-76) apply later.
// The output is bit-identical to oracle/ref_pack.py (tests/test_pack_host.py).
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "imp_b200.h"

namespace imp {
void set_error(const char* fmt, ...);
}

namespace {

template <class F>
void parallel_for(int64_t n, int n_threads, F&& fn) {
  if (n_threads <= 1 || n < 2048) {
    fn(0, n, 0);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=, &fn] { fn(lo, hi, t); });
  }
  for (auto& x : th) x.join();
}

struct MolView {
  const imp_ions_t* ions;
  int32_t idx;
};

// live entries of one ion as sortable keys: dst_local << 40 | bond << 24 | src_local
inline int build_keys(const imp_ions_t* ions, int32_t i, int32_t flags, int32_t max_edges, int32_t bond_vocab,
                      std::vector<uint64_t>& keys, int* err) {
  keys.clear();
  const int32_t n = ions->atom_ptr[i + 1] - ions->atom_ptr[i];
  const int32_t e0 = ions->edge_ptr[i], e1 = ions->edge_ptr[i + 1];
  const bool dbl = flags & IMP_PACK_DOUBLE_EDGES;
  const int32_t shift = (flags & IMP_PACK_SHIFT_IDS) ? 1 : 0;
  int64_t limit = max_edges >= 0 ? 2LL * max_edges : INT64_MAX;  // entries kept after the doubling
  int64_t produced = 0;
  for (int32_t e = e0; e < e1 && produced < limit; ++e) {
    const int32_t s = ions->edge_src[e], t = ions->edge_dst[e], b = ions->bond_ids[e] + shift;
    for (int rev = 0; rev < (dbl ? 2 : 1) && produced < limit; ++rev, ++produced) {
      const int32_t src = rev ? t : s, dst = rev ? s : t;
      if (src > 0 && dst > 0) {
        if (src >= n || dst >= n) { *err = IMP_ERR_INDEX; return 0; }
        if (b <= 0 || b >= bond_vocab) { *err = IMP_ERR_INDEX; return 0; }
        keys.push_back(((uint64_t)dst << 40) | ((uint64_t)b << 24) | (uint64_t)src);
      } else if (src < 0 || dst < 0) {
        *err = IMP_ERR_INDEX;
        return 0;
      }
    }
  }
  std::sort(keys.begin(), keys.end());
  int uniq = 0;
  for (size_t k = 0; k < keys.size(); ++k)
    if (k == 0 || keys[k] != keys[k - 1]) ++uniq;
  return uniq;
}

}  // namespace

extern "C" int imp_pack_host(const imp_ions_t* cation, const imp_ions_t* anion, int32_t bond_vocab, int32_t max_edges,
                             int32_t flags, int32_t edge_capacity, imp_graph_t* out, int32_t n_threads) {
  if (!cation || !anion || !out) { imp::set_error("imp_pack_host: null argument"); return IMP_ERR_ARG; }
  if (cation->n_ions != anion->n_ions || cation->n_ions < 0 || bond_vocab <= 0 || bond_vocab > 0xFFFF) {
    imp::set_error("imp_pack_host: n_ions mismatch or bad bond_vocab");
    return IMP_ERR_ARG;
  }
  const int32_t P = cation->n_ions, M = 2 * P;
  if (!out->mol_ptr || !out->row_ptr || !out->bucket_ptr || (edge_capacity > 0 && (!out->col_src || !out->edge_bm || !out->bucket_perm))) {
    imp::set_error("imp_pack_host: output arrays missing");
    return IMP_ERR_ARG;
  }
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  if (n_threads > 64) n_threads = 64;
  auto mol = [&](int32_t m) { return MolView{m < P ? cation : anion, m < P ? m : m - P}; };

  // atoms
  int64_t N = 0;
  out->mol_ptr[0] = 0;
  for (int32_t m = 0; m < M; ++m) {
    MolView v = mol(m);
    const int32_t n = v.ions->atom_ptr[v.idx + 1] - v.ions->atom_ptr[v.idx];
    if (n < 0 || n >= (1 << 24)) { imp::set_error("imp_pack_host: ion %d has %d atoms", m, n); return IMP_ERR_ARG; }
    N += n;
    if (N > INT32_MAX - 1) { imp::set_error("imp_pack_host: more than 2^31 atoms"); return IMP_ERR_CAPACITY; }
    out->mol_ptr[m + 1] = (int32_t)N;
  }
  if (N > 0 && !out->atom_id) { imp::set_error("imp_pack_host: atom_id missing"); return IMP_ERR_ARG; }
  const int32_t shift = (flags & IMP_PACK_SHIFT_IDS) ? 1 : 0;
  parallel_for(M, n_threads, [&](int64_t lo, int64_t hi, int) {
    for (int64_t m = lo; m < hi; ++m) {
      MolView v = mol((int32_t)m);
      const int32_t a0 = v.ions->atom_ptr[v.idx], n = v.ions->atom_ptr[v.idx + 1] - a0;
      for (int32_t k = 0; k < n; ++k) out->atom_id[out->mol_ptr[m] + k] = v.ions->atom_ids[a0 + k] + shift;
    }
  });

  // pass 1: unique live entries per molecule
  std::vector<int64_t> uniq_ptr(M + 1, 0);
  std::atomic<int> err{0};
  parallel_for(M, n_threads, [&](int64_t lo, int64_t hi, int) {
    std::vector<uint64_t> keys;
    for (int64_t m = lo; m < hi; ++m) {
      MolView v = mol((int32_t)m);
      int e = 0;
      uniq_ptr[m + 1] = build_keys(v.ions, v.idx, flags, max_edges, bond_vocab, keys, &e);
      if (e) err.store(e);
    }
  });
  if (err.load()) { imp::set_error("imp_pack_host: edge endpoint or bond id out of range"); return err.load(); }
  for (int32_t m = 0; m < M; ++m) uniq_ptr[m + 1] += uniq_ptr[m];
  const int64_t Eu = uniq_ptr[M];
  if (Eu > edge_capacity) {
    imp::set_error("imp_pack_host: %lld unique entries exceed edge_capacity %d", (long long)Eu, edge_capacity);
    return IMP_ERR_CAPACITY;
  }

  // pass 2: fill CSR
  std::vector<int64_t> edges_per_chunk(n_threads + 1, 0);
  parallel_for(M, n_threads, [&](int64_t lo, int64_t hi, int t) {
    std::vector<uint64_t> keys;
    int64_t e_mult = 0;
    for (int64_t m = lo; m < hi; ++m) {
      MolView v = mol((int32_t)m);
      int e = 0;
      build_keys(v.ions, v.idx, flags, max_edges, bond_vocab, keys, &e);
      const int32_t base = out->mol_ptr[m], n = out->mol_ptr[m + 1] - base;
      int64_t w = uniq_ptr[m];
      int32_t row = 0;
      out->row_ptr[base] = (int32_t)w;
      for (size_t k = 0; k < keys.size();) {
        size_t j = k;
        while (j < keys.size() && keys[j] == keys[k]) ++j;
        const int32_t dst = (int32_t)(keys[k] >> 40), b = (int32_t)((keys[k] >> 24) & 0xFFFF),
                      src = (int32_t)(keys[k] & 0xFFFFFF);
        int32_t mult = (int32_t)(j - k);
        if (mult > 0x7FFF) mult = 0x7FFF;  // cannot happen for featurize output; keeps the field well-formed
        while (row < dst) out->row_ptr[base + ++row] = (int32_t)w;
        out->col_src[w] = base + src;
        out->edge_bm[w] = b | (mult << 16);
        e_mult += mult;
        ++w;
        k = j;
      }
      while (row < n - 1) out->row_ptr[base + ++row] = (int32_t)w;
    }
    edges_per_chunk[t + 1] = e_mult;
  });
  out->row_ptr[N] = (int32_t)Eu;
  int64_t E = 0;
  for (int t = 1; t <= n_threads; ++t) E += edges_per_chunk[t];

  // buckets: stable counting sort of entry indices by (tower, bond)
  const int32_t G = 2 * bond_vocab;
  const int32_t n_cat = out->mol_ptr[P];
  const int64_t e_cat = N > 0 ? out->row_ptr[n_cat] : 0;  // entries [0, e_cat) have a cation destination
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, Eu / 65536 + 1));
  std::vector<int64_t> hist((size_t)nt * G, 0);
  const int64_t chunk = (Eu + nt - 1) / std::max(nt, 1);
  auto key_of = [&](int64_t e) { return (e >= e_cat ? bond_vocab : 0) + (out->edge_bm[e] & 0xFFFF); };
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const int64_t lo = t * chunk, hi = std::min(Eu, lo + chunk);
        for (int64_t e = lo; e < hi; ++e) hist[(size_t)t * G + key_of(e)]++;
      });
    for (auto& x : th) x.join();
  }
  int64_t run = 0;
  out->bucket_ptr[0] = 0;
  for (int32_t g = 0; g < G; ++g) {
    for (int t = 0; t < nt; ++t) {
      const int64_t c = hist[(size_t)t * G + g];
      hist[(size_t)t * G + g] = run;
      run += c;
    }
    out->bucket_ptr[g + 1] = (int32_t)run;
  }
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const int64_t lo = t * chunk, hi = std::min(Eu, lo + chunk);
        for (int64_t e = lo; e < hi; ++e) out->bucket_perm[hist[(size_t)t * G + key_of(e)]++] = (int32_t)e;
      });
    for (auto& x : th) x.join();
  }
  out->n_pairs = P;
  out->n_atoms = (int32_t)N;
  out->n_cat_atoms = n_cat;
  out->n_unique = (int32_t)Eu;
  out->n_edges = (int32_t)std::min<int64_t>(E, INT32_MAX);
  out->bond_vocab = bond_vocab;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Synthetic ions (SURVEY 8d recipe; same rules as ionic_mpnn_b200/synth.py, splitmix64 per ion so that
// generation is order-independent and parallel).
namespace {
struct Rng {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct IonShape {
  int32_t n, bonds;
};

// Generates one ion; if outputs are null only counts.  Returns {n_atoms, n_bonds}.
IonShape gen_ion(uint64_t seed, int32_t i, int32_t n_min, int32_t n_max, int32_t atom_types, int32_t bond_types,
                 const std::vector<double>* zipf_cdf, int32_t* atom_ids, int32_t* esrc, int32_t* edst, int32_t* bond) {
  Rng r{seed * 0xD1342543DE82EF95ull + (uint64_t)i * 0x2545F4914F6CDD1Dull + 1};
  r.next();
  const int32_t n = n_min + (int32_t)r.below((uint32_t)(n_max - n_min + 1));
  int32_t deg[4096];
  int32_t ba[4096 + 2], bb[4096 + 2];
  const int32_t nn = std::min(n, 4096);
  for (int32_t k = 0; k < nn; ++k) deg[k] = 0;
  int32_t nb = 0;
  for (int32_t k = 1; k < nn; ++k) {
    int32_t cand[4], nc = 0;
    for (int32_t j = std::max(0, k - 3); j < k; ++j)
      if (deg[j] < 4) cand[nc++] = j;
    int32_t j;
    if (nc) {
      j = cand[r.below((uint32_t)nc)];
    } else {
      int32_t cnt = 0;
      for (int32_t q = 0; q < k; ++q) cnt += deg[q] < 4;
      int32_t pick = (int32_t)r.below((uint32_t)cnt);
      j = 0;
      for (int32_t q = 0; q < k; ++q)
        if (deg[q] < 4 && pick-- == 0) { j = q; break; }
    }
    ba[nb] = j, bb[nb] = k, ++nb;
    deg[j]++, deg[k]++;
  }
  const int32_t rings = (int32_t)r.below(3);
  for (int32_t q = 0; q < rings; ++q) {
    int32_t a = (int32_t)r.below((uint32_t)nn), b = (int32_t)r.below((uint32_t)nn);
    int32_t lo = std::min(a, b), hi = std::max(a, b);
    if (lo == hi || deg[lo] >= 4 || deg[hi] >= 4) continue;
    bool adj = false;
    for (int32_t k = 0; k < nb; ++k)
      if (ba[k] == lo && bb[k] == hi) { adj = true; break; }
    if (adj) continue;
    ba[nb] = lo, bb[nb] = hi, ++nb;
    deg[lo]++, deg[hi]++;
  }
  if (atom_ids) {
    for (int32_t k = 0; k < nn; ++k) atom_ids[k] = (int32_t)r.below((uint32_t)atom_types);
    for (int32_t k = 0; k < nb; ++k) {
      int32_t t;
      if (zipf_cdf) {
        const double u = r.unit();
        t = (int32_t)(std::lower_bound(zipf_cdf->begin(), zipf_cdf->end(), u) - zipf_cdf->begin());
        t = std::min(t, bond_types - 1);
      } else {
        t = (int32_t)r.below((uint32_t)bond_types);
      }
      esrc[2 * k] = ba[k], edst[2 * k] = bb[k], bond[2 * k] = t;           // (a,b)
      esrc[2 * k + 1] = bb[k], edst[2 * k + 1] = ba[k], bond[2 * k + 1] = t;  // (b,a), src/featurize.py:60-63
    }
  }
  return {nn, nb};
}
}  // namespace

extern "C" int imp_synth_ions(uint64_t seed, int32_t n_ions, int32_t n_min, int32_t n_max, int32_t atom_types,
                              int32_t bond_types, int32_t skewed, int32_t* atom_ptr, int32_t* atom_ids, int32_t* edge_ptr,
                              int32_t* edge_src, int32_t* edge_dst, int32_t* bond_ids, int64_t* n_atoms,
                              int64_t* n_entries) {
  if (n_ions < 0 || n_min < 1 || n_max < n_min || n_max > 4096 || atom_types < 1 || bond_types < 1 || !n_atoms || !n_entries) {
    imp::set_error("imp_synth_ions: bad arguments");
    return IMP_ERR_ARG;
  }
  std::vector<double> cdf;
  if (skewed) {
    double tot = 0;
    for (int k = 1; k <= bond_types; ++k) tot += 1.0 / std::pow((double)k, 1.2);
    double run = 0;
    for (int k = 1; k <= bond_types; ++k) cdf.push_back(run += 1.0 / std::pow((double)k, 1.2) / tot);
  }
  const int nt = (int)std::min(64u, std::max(1u, std::thread::hardware_concurrency()));
  if (!atom_ptr) {  // phase 1: counts only
    std::vector<int64_t> na(nt, 0), ne(nt, 0);
    parallel_for(n_ions, nt, [&](int64_t lo, int64_t hi, int t) {
      for (int64_t i = lo; i < hi; ++i) {
        IonShape s = gen_ion(seed, (int32_t)i, n_min, n_max, atom_types, bond_types, nullptr, nullptr, nullptr, nullptr, nullptr);
        na[t] += s.n, ne[t] += 2 * s.bonds;
      }
    });
    *n_atoms = 0, *n_entries = 0;
    for (int t = 0; t < nt; ++t) *n_atoms += na[t], *n_entries += ne[t];
    return 0;
  }
  if (!atom_ids || !edge_ptr || !edge_src || !edge_dst || !bond_ids) {
    imp::set_error("imp_synth_ions: output arrays missing");
    return IMP_ERR_ARG;
  }
  std::vector<IonShape> shapes(n_ions);
  parallel_for(n_ions, nt, [&](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi; ++i)
      shapes[i] = gen_ion(seed, (int32_t)i, n_min, n_max, atom_types, bond_types, nullptr, nullptr, nullptr, nullptr, nullptr);
  });
  int64_t a = 0, e = 0;
  atom_ptr[0] = 0, edge_ptr[0] = 0;
  for (int32_t i = 0; i < n_ions; ++i) {
    a += shapes[i].n, e += 2 * shapes[i].bonds;
    if (a > INT32_MAX || e > INT32_MAX) { imp::set_error("imp_synth_ions: batch too large for int32 offsets"); return IMP_ERR_CAPACITY; }
    atom_ptr[i + 1] = (int32_t)a, edge_ptr[i + 1] = (int32_t)e;
  }
  parallel_for(n_ions, nt, [&](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi; ++i)
      gen_ion(seed, (int32_t)i, n_min, n_max, atom_types, bond_types, skewed ? &cdf : nullptr, atom_ids + atom_ptr[i],
              edge_src + edge_ptr[i], edge_dst + edge_ptr[i], bond_ids + edge_ptr[i]);
  });
  *n_atoms = a, *n_entries = e;
  return 0;
}
