// Tile plan of the fused whole-tower forward (fifth generation): the packed graph batch re-cut into self-contained
// 128-row tiles, one fixed-size record per tile, so that the forward kernel's per-tile prologue is ONE TMA bulk copy.
//
// A record holds whole molecules of one tower (no edge crosses a molecule: models/layers.py:100-117 gathers inside one
// padded ion).  Molecules are chosen best-fit from a window of FP_WIN consecutive molecules, not contiguously, so tiles
// are ~97 % full (the contiguous greedy cut of generations 1-4 leaves ~12 % of the rows empty); the rows are listed in
// order of in-degree (the row -> thread assignment that keeps the 32 lanes of a warp on the same entry count) and the
// CSR entries are already translated to tile-local rows.  Built by imp_fused_plan (fused_plan.cu) once per batch from
// either input feed, consumed by mpnn_fused_h5_kernel (fused_fwd5.cu).
#pragma once
#include <stdint.h>

namespace imp {

constexpr int FP_ROWS = 128;
constexpr int FP_ECAP = 336;    // entries per tile (natural row order)
constexpr int FP_MAXMOL = 32;   // molecules per tile
constexpr int FP_WIN = 256;     // molecules per packing window
constexpr int FP_HEADER_BYTES = 256;

struct alignas(16) FusedTile {
  uint32_t slot[FP_ROWS];         // sorted position -> natural row | in-degree << 7 | first entry << 12 | atom id << 22
  uint32_t ent[FP_ECAP];          // src (natural row) | bond << 8 | multiplicity (IEEE half bits) << 16
  int32_t molid[FP_MAXMOL];       // row of d_pooled (tower-major molecule index)
  uint8_t mol_lo[FP_MAXMOL + 4];  // first natural row of molecule j; mol_lo[nm] = rows
  uint8_t nm, rows;
  uint16_t n_ent;
  uint8_t pad[24];
};
static_assert(sizeof(FusedTile) == 2048, "tile record is one 2 KiB TMA bulk copy");

struct FusedPlanHeader {  // first FP_HEADER_BYTES of the plan buffer, then cap[0] cation tiles, then cap[1] anion tiles
  int32_t n_tiles[2];
  int32_t cap[2];
  int32_t status;  // 0, or 1: a molecule outside the envelope (> 128 atoms, a row with > 31 entries, > FP_ECAP entries),
                   //        2: tile capacity exceeded
  int32_t pad[3];  // [0], [1]: per-tower tile tickets of the sixth-generation forward (zeroed by its launcher)
};

}  // namespace imp
