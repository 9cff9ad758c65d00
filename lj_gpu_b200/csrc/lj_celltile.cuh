// lj_celltile.cuh -- geometry shared by the builder (lj_nlist.cu) and the consumer
// (lj_force_celltile.cu) of the cell-tile mirror.
//
// The list build sorts particles by cell (x fastest).  A TILE (tx, cy, cz) is a run of `tc`
// x-consecutive cells of the pencil (cy, cz); its rows are ONE contiguous range of the cell order.
// Every neighbour of a tile row lies in one of 25 pencils (cy+dy-2, cz+dz-2), x-cells
// [xa-2, xb+2].  The five pencils that share a y form a Y-ROW (tx, Y, cz): five contiguous ranges
// of the cell-ordered position array.  Tile cy needs y-rows cy-2 .. cy+2, tile cy+1 needs
// cy-1 .. cy+3: a CTA that walks a COLUMN (tx, cz) in y keeps a ring of y-rows in shared memory
// and stages ONE new y-row per tile (five TMA bulk copies) instead of the whole 25-pencil region.
//
// Region-local index of a particle of pencil (dy, dz), cell-order index m:
//     L = dy * cap_y + pb[Y][dz] + (m - st[Y][dz]),   Y = cy + dy - 2
// with {st, pb} from the y-row table and cap_y = the longest y-row + 8 (one common stride, so the
// index does not depend on where the ring currently holds the row).  The last record of each ring
// slot is never written by a copy and holds a far-away point: L = cap_y - 1 is the DUMMY index rows
// are padded with.  16-bit entries: 5 * cap_y must stay below 65536.
#pragma once
#include "lj_common.cuh"

constexpr int kTileYPencils = 5;   // pencils per y-row
constexpr int kTileYTab = 6;       // uint2 per y-row in the table: 5 x {st, pb}, then {0, length}
constexpr int kTileTTab = 2;       // uint4 per tile: {s0, rows, u0, units}, {self0, 0, 0, 0}
constexpr double kTileFar = 1.0e10;  // coordinates of the dummy record

// x-extent of the tile that holds cell column cx, and of its region
__device__ __forceinline__ void tile_x_extent(int cx, int tc, int nx, int& xa, int& xb, int& rxa,
                                              int& rxb) {
  xa = (cx / tc) * tc;
  xb = min(xa + tc - 1, nx - 1);
  rxa = max(xa - 2, 0);
  rxb = min(xb + 2, nx - 1);
}

// Pencil dz of the y-row (Y, cz): cell-order range [st, st+len), both ends moved outwards to even
// indices (16-byte granularity of the packed double3 records).  len = 0 outside the grid.
__device__ __forceinline__ void tile_pencil_range(const uint32_t* __restrict__ cell_start, int nx,
                                                  int ny, int nz, int Y, int cz, int rxa, int rxb,
                                                  int dz, uint32_t& st, uint32_t& len) {
  const int z = cz + dz - 2;
  st = 0; len = 0;
  if (Y < 0 || Y >= ny || z < 0 || z >= nz) return;
  const int rowc = (z * ny + Y) * nx;
  st = cell_start[rowc + rxa] & ~1u;
  len = ((cell_start[rowc + rxb + 1] + 1u) & ~1u) - st;
}
