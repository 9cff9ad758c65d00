// lj_force_celltile.cu -- force kernels on the cell-tile mirror (sm_100a): FP64 and mixed precision.
//
// Why: the per-row gather kernels are bound by the L1/LSU data pipe.  A 32-lane LDG.E.256 gather
// costs ~2 SM-cycles per distinct 128-byte line it touches (8-32 lines), the 4-byte list words are
// a second LSU stream, and both wait on L2/DRAM latency.  This kernel takes the gather off the
// global path altogether (geometry: lj_celltile.cuh):
//
//   * persistent CTAs, one per SM.  A CTA walks COLUMNS of tiles (fixed x-range and z, y
//     ascending).  The neighbours of a tile are the 25 pencils (y-2..y+2) x (z-2..z+2); the five
//     pencils of one y form a y-row, and consecutive tiles share four of their five y-rows.  Two
//     producer warps keep the rings in shared memory filled: warp Y stages ONE new y-row per tile
//     (five TMA bulk copies, cp.async.bulk -> UBLKCP; ten for FP64, see below), warp L the tile's
//     list segment, its row metadata and its header; completions are counted in bytes on the
//     tile's mbarrier: ~16 KB per tile instead of the 60 KB of a whole region;
//   * the mirror list holds 16-bit region-local indices (2 B per pair from HBM instead of 4),
//     rows padded to 8 entries with an index that points at a far-away dummy point, so the inner
//     loop has no bounds logic;
//   * FP64: the ring holds a plane of {x,y} pairs and a plane of z; per pair-iteration a warp
//     issues LDS.U16 (index), LDS.128 and LDS.64 at ~30-cycle latency instead of one list LDG and
//     one 32-line gather at L2 latency;
//   * mixed precision (LJ_PREC_MIXED): the ring holds 16-byte fixed-point records {x, y, z,
//     original index}, 32-bit counts of a power-of-two unit modulo 2^32 (lj_fx_frame): one LDS.128
//     per pair, exact integer differences, FP32 pair arithmetic (MUFU.RCP), FP32 per-lane partial
//     sums, FP64 reduction and momenta.  Pairs within the FP32 error band of the cutoff are left
//     out of the hot loop and re-decided in FP64 from the caller's positions (rare rows only).
//
// Twenty consumer warps (FP64; twenty-four in mixed precision: LJ_CT_NCONS / LJ_CT_NCONS_MX): eight lanes per row, four rows (a quad) per warp in lock step with
// warp-uniform trip counts, quads dealt round-robin across tiles, shuffle reduction, one
// RED.ADD.F64 per component and row (exactly one add per step: deterministic).  FP64 results are
// bit-identical to the per-row kernel with group = 8 on the same list order.  Positions are
// re-permuted into cell order at the start of every step (k_tile_permute[_fx], ~10 us at N = 1M),
// so moving particles are handled exactly like in the per-row kernels: the list decides
// membership, the current q decides the force.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "lj_celltile.cuh"
#include "lj_tile.cuh"

namespace {

#ifndef LJ_CT_UNROLL
#define LJ_CT_UNROLL 4
#endif
constexpr int kCtUnroll = LJ_CT_UNROLL;
#ifndef LJ_CT_UNROLL_MX
#define LJ_CT_UNROLL_MX 4
#endif
constexpr int kCtUnrollMx = LJ_CT_UNROLL_MX;
#ifndef LJ_CT_LANES_MX
#define LJ_CT_LANES_MX 8
#endif
constexpr int kCtLanesMx = LJ_CT_LANES_MX;  // lanes per row in the mixed kernel: 8 or 4
#ifndef LJ_CT_LIST_DEPTH
#define LJ_CT_LIST_DEPTH 5
#endif
#ifndef LJ_CT_GRADED
#define LJ_CT_GRADED 1
#endif
#ifndef LJ_CT_GRADE_PCT
#define LJ_CT_GRADE_PCT 45
#endif
#ifndef LJ_CT_GRADE_MIN
#define LJ_CT_GRADE_MIN 3
#endif
constexpr bool kCtGraded = LJ_CT_GRADED != 0;  // graded column segments (see launch_celltile)
// Consumer warps of the product kernels.  Measured on the byte-granular ring (same box, 100 launches, bit-exact):
// FP64 16 / 17 / 18 / 20 / 22 / 24 warps -> 0.2926 / 0.3163 / 0.3024 / 0.2904 / 0.2974 / 0.2927 ms (80 registers at 20,
// no spills); mixed 16 / 20 / 22 / 24 -> 0.2343 / 0.2316 / 0.2305 / 0.2284 ms (64-72 registers).
#ifndef LJ_CT_NCONS
#define LJ_CT_NCONS 20
#endif
#ifndef LJ_CT_NCONS_MX
#define LJ_CT_NCONS_MX 24
#endif
#ifndef LJ_CT_ALU_SUB
#define LJ_CT_ALU_SUB 0  // 1: integer differences as VIADDMNMX on the ALU pipe (measured: no gain)
#endif
#ifndef LJ_CT_DIAG
#define LJ_CT_DIAG 0  // 1: the mixed kernel honours LJ_TILE_MODE = 1 (no pair math) and 2 (no gather)
#endif

struct __align__(16) tile_hdr { int ns, self0; uint32_t u0; int yslot0; };

template <int LAYOUT>
__device__ __forceinline__ void red_mom(void* __restrict__ p, int64_t i, int64_t plane, double fx,
                                        double fy, double fz) {
  double* b;
  int64_t s;
  if (LAYOUT == LJ_AOS_D4) { b = reinterpret_cast<double*>(p) + 4 * i; s = 1; }
  else if (LAYOUT == LJ_AOS_D3) { b = reinterpret_cast<double*>(p) + 3 * i; s = 1; }
  else { b = reinterpret_cast<double*>(p) + i; s = plane; }
  atomicAdd(b, fx);
  atomicAdd(b + s, fy);
  atomicAdd(b + 2 * s, fz);
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

#ifndef LJ_CT_MAXY
#define LJ_CT_MAXY 16
#endif
constexpr int kCtMaxY = LJ_CT_MAXY, kCtMaxL = 8;
constexpr int kCtMaxSeg = 32;  // tiles per unit (column segment) at most
constexpr int kCtMaxUnitsPerCol = 32;  // segments a column is cut into at most

struct ct_params {
  const unsigned char* qs;  // positions in cell order: int4 fixed point (mixed) or the {x,y} plane (FP64)
  const unsigned char* qz;  // FP64: the z plane
  void* p; int64_t plane;
  double c24, c48; long long cl2_bits;
  // mixed precision only
  const void* q;            // the caller's FP64 positions: exact cutoff decision of borderline pairs
  const lj_grid_params* grid;
  double cl2, fx_scale;     // counts per length (a power of two)
  float c24u, c48u;         // 24 dt unit, 48 dt unit (the differences stay in counts)
  float unit2, cl2f, lo_c;  // unit^2; the cutoff^2; r2 <= lo_c = cl2f - margin: inside for sure
  float band;               // |r2 - cl2f| <= band: too close to call in FP32 (slightly above margin)
  const uint2* ytab; const uint4* ttab; const int4* meta; const uint16_t* list;
  int ntx, ny, ncols;    // tiles per pencil, cells in y, ACTIVE columns (all of them: ntx * nz)
  const int32_t* cols;   // the active columns, or NULL when every column is active
  const int* ncols_dev;  // part launches: the number of selected columns is only known on the device
  int seg_len, nseg;     // a unit = one column x the tiles [seg_y0[seg], seg_y0[seg + 1]) (seg_len: the longest)
  short seg_y0[kCtMaxUnitsPerCol + 1];
  int cap_y, cap_units, cap_rows;
  int ry, rl;            // ring sizes: y-row slots, tiles in flight (headers + barriers)
  int lring_bytes;       // the list ring: every tile takes exactly what its list and row metadata need
  int* unit_counter;     // zeroed by k_tile_permute before every launch: units are dealt dynamically
  long long* dbg;        // diagnostics (LJ_TILE_DBG): per consumer warp {wait, work, quads, total} cycles
  int mode;              // diagnostics (LJ_TILE_MODE): 0 normal, 1 no pair math, 3 staging only
};

// One `full` and one `empty` mbarrier per TILE slot.  Everything a tile waits for -- its newest
// y-row (all five for the first tile of a column segment), its list segment and its row metadata --
// completes on the tile's full barrier, so a consumer warp pays one wait, one 16-byte header read
// and one arrive per tile (most warps have no quad in a given tile: that path must be short).  The
// y-row ring is managed by the producer alone: a y-row's slot is free once the tile that had it as
// its oldest row has been released.
//
// MX = mixed precision: the ring holds 16-byte fixed-point records {x, y, z, original index}
// (one LDS.128 per pair instead of three LDS.64 on 24-byte records), differences are exact integer
// subtractions, the pair arithmetic is FP32 and the per-row sums are reduced and added to p in
// FP64 -- see the consumer loop.  Producer, rings and barriers are the same for both precisions.
// NB = CTAs per SM.  The producer warp needs ~1600 cycles of dependent instructions per tile
// (measured: 0.14 ms per step with every copy and all pair work switched off), which is half of
// what 16 consumer warps need to work a tile off -- any hiccup and they wait (22 % of their time).
// Two CTAs of 8 consumer warps per SM are two independent pipelines at half the tile rate each.
// Every consumer warp waits on EVERY tile's full barrier, in order.  (Round 2 tried dealing the
// tiles to groups of warps so that a warp visits only every second or fourth tile: 2 % faster and
// WRONG -- a tile's older y-rows complete on the barriers of the preceding tiles, which a warp that
// skips those tiles never waits for; results were off by 1e-6..1e-5 in one run out of a few.)
template <int LAYOUT, bool MX, int NCONS, int NB>
__global__ void __launch_bounds__((NCONS + 2) * 32, NB)
lj_celltile_force(const ct_params P) {
#if LJ_DIAG  // per-warp cycle counters and kernel-surgery modes exist in diagnostic builds only
  long long* const dbgp = P.dbg;
  const int modev = P.mode;
#else
  constexpr long long* dbgp = nullptr;
  constexpr int modev = 0;
#endif
  constexpr uint32_t RB = MX ? 16u : 24u;  // bytes per staged position record
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t tfull[kCtMaxL], tempty[kCtMaxL];
  __shared__ tile_hdr hdr[kCtMaxL];
  __shared__ uint32_t lstart[kCtMaxL];  // warp L: first byte of the tile's data in the list ring
  __shared__ int yrel[kCtMaxY];  // producer: tile sequence number whose release frees the y slot
  // the producer's view of the y-row / tile tables of the current and the next unit (bulk-copied one
  // unit ahead: a table entry fetched with a plain load costs a DRAM round trip per tile)
  __shared__ __align__(16) uint2 ytab_s[2][(kCtMaxSeg + 4) * kTileYTab];
  __shared__ __align__(16) uint4 ttab_s[2][kCtMaxSeg * kTileTTab];
  __shared__ __align__(8) uint64_t tabbar[2][2];     // [warp Y | warp L][staging buffer]
  __shared__ __align__(8) uint64_t mfull[2], mempty[2];  // unit mailbox, warp Y -> warp L
  __shared__ int umail[2];
  __shared__ float kconst[8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ry = P.ry, rl = P.rl, cap_y = P.cap_y;
  // FP64 ring: a plane of {x,y} pairs (16 B, one LDS.128) followed by a plane of z (8 B, LDS.64), both
  // indexed by ring record.  Packed 24-byte records cost three LDS.64 whose half-warps (two rows
  // of eight lanes at unrelated offsets) almost always collide: 16 shared-memory wavefronts per 32
  // pairs measured, against 6 for conflict-free access; the planes need 4 + ~4.
  unsigned char* const ybase = smem_raw;
  unsigned char* const zbase = smem_raw + (size_t)ry * cap_y * 16;  // FP64 only
  unsigned char* const lbase = smem_raw + (size_t)ry * cap_y * RB;
  if (threadIdx.x == 0) {
    kconst[0] = P.unit2; kconst[1] = P.c24u; kconst[2] = P.c48u; kconst[3] = P.lo_c; kconst[4] = P.cl2f;
    kconst[5] = __int_as_float(0x7fffffff);
    for (int b = 0; b < rl; b++) { mbar_init(&tfull[b], 2); mbar_init(&tempty[b], NCONS); }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tabbar[0][b], 1); mbar_init(&tabbar[1][b], 1);
      mbar_init(&mfull[b], 1); mbar_init(&mempty[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < ry) {  // the dummy record of every y slot: its last one, no copy ever reaches it
    if (!MX) {
      const size_t R = (size_t)threadIdx.x * cap_y + cap_y - 1;
      reinterpret_cast<double2*>(ybase)[R] = make_double2(kTileFar, kTileFar);
      reinterpret_cast<double*>(zbase)[R] = kTileFar;
    }  // MX: fixed-point coordinates wrap, "far" depends on the unit -- the producer writes it
    yrel[threadIdx.x] = -1;
  }
  __syncthreads();
  const int ncols = P.ncols_dev ? *P.ncols_dev : P.ncols;  // (written by the permute kernel that ran before)
  const int nunits = ncols * P.nseg;

  if (warp >= NCONS) {
    // ------------------------------------------------------- two producer warps, by role ---
    // One warp needs ~2800 cycles of dependent instructions per tile (mbarrier polls, table reads,
    // shuffles, seven serialised UBLKCP) while the consumers work a tile off in ~2100-2700: measured
    // with LJ_TILE_DBG, the single producer was busy 93 % of the kernel and the consumers waited.
    // Warp Y stages the y-rows (and the mixed kernel's dummy record), warp L the list segment, the
    // row metadata and the header; each arrives once on the tile's full barrier (count 2).  Both
    // walk the same unit sequence: Y claims units from the global counter and posts them to L
    // through a two-entry mailbox.
    const bool isY = warp == NCONS;
    int tseq = 0, tslot = 0;             // tile being assembled: sequence number, slot (ring of rl)
    int done = 0, dslot = 0, dphase = 0; // tiles known to be released: [0, done)
    uint32_t lhead = 0u;                 // warp L: next free byte of the list ring
    const uint32_t lring = (uint32_t)P.lring_bytes;
    long long p_idle = 0;                // diagnostics: cycles spent waiting for the consumers
    const long long p_begin = dbgp ? clock64() : 0;
    auto ensure_done = [&](int q) {      // block until tile q has been released by every consumer
      while (done <= q) {
        const long long w0 = dbgp ? clock64() : 0;
        mbar_wait(&tempty[dslot], dphase);
        if (dbgp) p_idle += clock64() - w0;
        done++;
        if (++dslot == rl) { dslot = 0; dphase ^= 1; }
      }
    };
    // k-th unit of this CTA.  Units are dealt dynamically (tiles differ in size by 2x, a static deal
    // leaves SMs idle at the end); a unit is claimed one ahead so that its tables can be staged early.
    // The global atomic takes ~2000 cycles to return: warp Y issues the claim of unit k + 1 while it
    // takes delivery of unit k (the value sits in lane 0's register for a whole unit).
    int claim_pending = 0;
    if (isY && lane == 0) claim_pending = atomicAdd(P.unit_counter, 1);
    auto next_unit = [&](int k) {
      const int b = k & 1;
      int v = 0;
      if (isY) {
        v = __shfl_sync(0xffffffffu, claim_pending, 0);
        if (lane == 0 && v < nunits) claim_pending = atomicAdd(P.unit_counter, 1);  // claim k + 1, used a unit later
        if (k >= 2) mbar_wait(&mempty[b], ((k >> 1) - 1) & 1);  // L has read entry k - 2
        if (lane == 0) { umail[b] = v; mbar_arrive(&mfull[b]); }
      } else {
        mbar_wait(&mfull[b], (k >> 1) & 1);
        v = *reinterpret_cast<volatile int*>(&umail[b]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&mempty[b]);
      }
      return v;
    };
    // table rows of unit u -> staging buffer b: y-rows max(y0-2,0) .. min(y1+1,ny-1) (warp Y),
    // tiles y0 .. y1-1 (warp L)
    auto stage_tables = [&](int u, int b) {  // returns the unit's column (tx, cz)
      const int ci = u % ncols, seg = u / ncols;
      const int col = P.cols ? __ldg(P.cols + ci) : ci;
      const int y0 = P.seg_y0[seg], y1 = P.seg_y0[seg + 1];
      const int ylo = max(y0 - 2, 0), yhi = min(y1 + 1, P.ny - 1);
      if (lane == 0) {
        if (isY) {
          const uint32_t yb = (uint32_t)(yhi - ylo + 1) * kTileYTab * 8u;
          mbar_arrive_expect_tx(&tabbar[0][b], yb);
          bulk_g2s(&ytab_s[b][0], P.ytab + ((size_t)col * P.ny + ylo) * kTileYTab, yb, &tabbar[0][b]);
        } else {
          const uint32_t tb = (uint32_t)(y1 - y0) * kTileTTab * 16u;
          mbar_arrive_expect_tx(&tabbar[1][b], tb);
          bulk_g2s(&ttab_s[b][0], P.ttab + ((size_t)col * P.ny + y0) * kTileTTab, tb, &tabbar[1][b]);
        }
      }
      return col;
    };
    int yslot = 0;  // warp Y: next y-row slot (ring of ry)
    double grid_oz = 0.0, grid_edge = 0.0;  // MX, warp Y: read once (two global loads and a division)
    if (MX && isY) { grid_oz = P.grid->oz; grid_edge = 1.0 / P.grid->inv_cell; }
    int nu = 0;     // units done by this CTA
    int u = next_unit(0);
    int col_cur = 0, col_next = 0;
    if (u < nunits) col_cur = stage_tables(u, 0);
    for (; u < nunits; nu++) {
      const int u_next = next_unit(nu + 1);
      const int seg = u / ncols;
      const int y0 = P.seg_y0[seg], y1 = P.seg_y0[seg + 1];
      const int ntile = y1 - y0;
      const int tb = nu & 1;
      __syncwarp();  // every lane is done with the other buffer (the previous unit's tables)
      if (u_next < nunits) col_next = stage_tables(u_next, tb ^ 1);
      mbar_wait(&tabbar[isY ? 0 : 1][tb], (nu >> 1) & 1);
      if (isY) {
        // ================================================================ warp Y: y-rows
        // MX: the dummy record rows are padded with.  Fixed-point coordinates live modulo 2^32
        // counts, so no point is far from everything; half a period away in z from this unit's
        // cell layer is far from every row of the unit (all in z-cell cz, give or take the skin).
        int dummy_z = 0;
        if (MX) {
          const int cz = col_cur / P.ntx;
          const double zc = grid_oz + ((double)cz + 0.5) * grid_edge;
          dummy_z = (int)((uint32_t)__double2ll_rn(zc * P.fx_scale) + 0x80000000u);
        }
        const int ylo = max(y0 - 2, 0);
        auto load_y = [&](int Y) {  // lane dz < 5: {st, pb}; lane 5: {0, length}
          uint2 e = make_uint2(0u, 0u);
          if (lane < kTileYTab && Y >= 0 && Y < P.ny) e = ytab_s[tb][(Y - ylo) * kTileYTab + lane];
          return e;
        };
        uint2 ey_next = load_y(y0 - 2);
        const int tbase = tseq;
        int first_yslot = yslot;  // slot of the oldest y-row of the tile being assembled
        for (int i = 0; i < ntile + 4; i++) {
          const int Y = y0 - 2 + i;
          const uint2 ey = ey_next;
          ey_next = load_y(Y + 1);  // in flight while this iteration waits and issues
          // the tile this y-row completes on: tile 0 for the first five rows, then one row per tile.
          // Its slot (barrier, header) must be free before the first contribution.
          if (i == 0 || i > 4) { if (tseq >= rl) ensure_done(tseq - rl); }
          const int rel = yrel[yslot];
          if (rel >= 0) ensure_done(rel);
          const uint32_t pb_next = __shfl_down_sync(0xffffffffu, ey.y, 1);
          const uint32_t len = lane < kTileYPencils ? pb_next - ey.y : 0u;
          uint32_t ylen = __shfl_sync(0xffffffffu, ey.y, 5);
          if (modev & 16) ylen = 0;  // diagnostics: no y-row copies
          if ((int)ylen > cap_y - 8) __trap();
          if (lane == 0) {
            yrel[yslot] = tbase + min(i, ntile - 1);  // the tile that has it as its oldest row
            if (ylen) mbar_expect_tx(&tfull[tslot], ylen * RB);
            if (MX)  // made visible to the consumers by this lane's arrive on the tile's full barrier
              *reinterpret_cast<int4*>(ybase + ((size_t)yslot * cap_y + cap_y - 1) * RB) = make_int4(0, 0, dummy_z, 0);
          }
          __syncwarp();
          if (len && ylen) {
            const size_t R = (size_t)yslot * cap_y + ey.y;
            if (MX) {
              bulk_g2s(ybase + R * 16, P.qs + (size_t)ey.x * 16, len * 16u, &tfull[tslot]);
            } else {  // both ends of a pencil range are even: the 8-byte plane stays 16-byte aligned
              bulk_g2s(ybase + R * 16, P.qs + (size_t)ey.x * 16, len * 16u, &tfull[tslot]);
              bulk_g2s(zbase + R * 8, P.qz + (size_t)ey.x * 8, len * 8u, &tfull[tslot]);
            }
          }
          if (++yslot == ry) yslot = 0;
          if (i >= 4) {  // the tile's five y-rows are on their way: this warp's one arrival
            if (lane == 0) {
              hdr[tslot].yslot0 = first_yslot;
              mbar_arrive(&tfull[tslot]);
            }
            tseq++;
            if (++tslot == rl) tslot = 0;
            if (++first_yslot == ry) first_yslot = 0;
          }
        }
      } else {
        // ================================================================ warp L: list, metadata, header
        auto load_t = [&](int cy) {  // lanes 0, 1: the two uint4 of the tile
          uint4 e = make_uint4(0u, 0u, 0u, 0u);
          if (lane < kTileTTab && cy >= y0 && cy < y1) e = ttab_s[tb][(cy - y0) * kTileTTab + lane];
          return e;
        };
        uint4 et_next = load_t(y0);
        for (int cy = y0; cy < y1; cy++) {
          if (tseq >= rl) ensure_done(tseq - rl);
          const uint4 et = et_next;
          et_next = load_t(cy + 1);
          const uint32_t s0 = __shfl_sync(0xffffffffu, et.x, 0), ns = __shfl_sync(0xffffffffu, et.y, 0);
          const uint32_t u0 = __shfl_sync(0xffffffffu, et.z, 0), units = __shfl_sync(0xffffffffu, et.w, 0);
          const uint32_t self0 = __shfl_sync(0xffffffffu, et.x, 1);
          const uint32_t units_c = (modev & 32) ? 0u : units, ns_c = (modev & 64) ? 0u : ns;  // diagnostics
          if ((int)units > P.cap_units || (int)ns > P.cap_rows) __trap();
          // The list ring is BYTE-granular: a tile takes [row metadata | list units] = (ns + units) * 16 bytes,
          // contiguous (a bulk copy does not wrap), wherever the ring has room.  Slots sized for the longest
          // tile were half empty on average (tiles hold 30-70 rows), and what the ring saves went to y-row
          // slots.  Tiles are released in order, so the free space is what lies between the head and the
          // first byte of the oldest tile still in use.
          const uint32_t need = (ns + units) * 16u;
          for (;;) {
            if (done == tseq) { lhead = 0u; break; }  // nothing in flight: start over at the bottom
            const uint32_t tail = *reinterpret_cast<volatile uint32_t*>(&lstart[dslot]);
            if (lhead >= tail) {
              if (lhead + need <= lring) break;             // room above the head
              if (need < tail) { lhead = 0u; break; }       // wrap: room below the oldest tile
            } else if (lhead + need < tail) break;          // wrapped already: room up to the oldest tile
            ensure_done(done);                              // wait for the oldest tile, look again
          }
          unsigned char* dst = lbase + lhead;
          if (lane == 0) {
            lstart[tslot] = lhead;
            hdr[tslot].ns = (int)(ns | (lhead >> 4) << 16);  // low half: rows, high half: ring offset / 16
            hdr[tslot].self0 = (int)self0 + 2 * cap_y; hdr[tslot].u0 = u0;
            mbar_arrive_expect_tx(&tfull[tslot], units_c * 16u + ns_c * 16u);
          }
          __syncwarp();
          if (lane == 1 && ns_c) bulk_g2s(dst, P.meta + s0, ns * 16u, &tfull[tslot]);
          if (lane == 0 && units_c) bulk_g2s(dst + (size_t)ns * 16, P.list + (size_t)u0 * 8, units * 16u, &tfull[tslot]);
          lhead += need;
          tseq++;
          if (++tslot == rl) tslot = 0;
        }
      }
      u = u_next;
      col_cur = col_next;
    }
    // end marker for the consumers: a tile header with ns = 0xffff (both producer warps arrive)
    if (tseq >= rl) ensure_done(tseq - rl);
    if (lane == 0) {
      if (!isY) hdr[tslot].ns = 0xffff;
      mbar_arrive(&tfull[tslot]);
      if (dbgp) {  // producer records: {idle, 0, tiles, total}
        long long* d = dbgp + ((size_t)gridDim.x * NCONS + 2 * blockIdx.x + (isY ? 0 : 1)) * 4;
        d[0] = p_idle; d[1] = 0; d[2] = tseq; d[3] = clock64() - p_begin;
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps ---
  // MX: the pair-loop constants, read back through a volatile shared-memory load so that they
  // live in registers.  Left as kernel parameters, ptxas re-reads the constant bank inside the
  // pair loop (9 of 116 instructions per four pairs).
  float unit2 = 0.f, c24u = 0.f, c48u = 0.f, lo_c = 0.f, cl2f = 0.f;
  int kbig = 0x7fffffff;  // INT_MAX, opaque to ptxas (LJ_CT_ALU_SUB)
  if (MX) {
    asm volatile("ld.volatile.shared.s32 %0, [%1+20];" : "=r"(kbig) : "r"(smem_u32(kconst)));
    const uint32_t ka = smem_u32(kconst);
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(unit2) : "r"(ka));
    asm volatile("ld.volatile.shared.f32 %0, [%1+4];" : "=f"(c24u) : "r"(ka));
    asm volatile("ld.volatile.shared.f32 %0, [%1+8];" : "=f"(c48u) : "r"(ka));
    asm volatile("ld.volatile.shared.f32 %0, [%1+12];" : "=f"(lo_c) : "r"(ka));
    asm volatile("ld.volatile.shared.f32 %0, [%1+16];" : "=f"(cl2f) : "r"(ka));
  }
  const int lg = lane & 7, gi = lane >> 3;
  const uint32_t ring = (uint32_t)ry * (uint32_t)cap_y;
  const uint32_t ybase_s = smem_u32(ybase);
  const uint32_t dummy = (uint32_t)cap_y - 1u;
  int tslot = 0, tphase = 0;
  long long t_wait = 0, t_work = 0, n_quads = 0;
  const long long t_begin = dbgp ? clock64() : 0;
  int first = warp;  // quads are dealt round-robin over the warps ACROSS tiles: tiles hold fewer
                     // quads than there are warps, a per-tile deal would leave the high warps idle
  for (;;) {
    {
      long long tw0 = 0;
      if (dbgp) tw0 = clock64();
      mbar_wait(&tfull[tslot], tphase);
      long long tw1 = 0;
      if (dbgp) { tw1 = clock64(); t_wait += tw1 - tw0; }
      const int4 h = *reinterpret_cast<const int4*>(&hdr[tslot]);  // ns, self0, u0, yslot0
      const int ns = h.x & 0xffff;
      if (ns == 0xffff) break;  // end marker
      constexpr int kRows = MX ? 32 / kCtLanesMx : 4;  // rows a warp works on in lock step
      const int nquads = (ns + kRows - 1) / kRows;
      int quad = first;
      const bool had_quad = quad < nquads;
      first = (first + NCONS - nquads % NCONS) % NCONS;
      if (quad < nquads && (modev & 15) != 3) {
        const int self0 = h.y;
        const uint32_t u0 = (uint32_t)h.z;
        const unsigned char* lptr = lbase + (size_t)((uint32_t)h.x >> 16) * 16;  // [row metadata | list units]
        const int4* __restrict__ meta = reinterpret_cast<const int4*>(lptr);
        const uint16_t* __restrict__ lst = reinterpret_cast<const uint16_t*>(lptr + (size_t)ns * 16);
        // region-local index L -> ring record: (slot of the tile's first y-row) * cap_y + L, wrapped
        const uint32_t off0 = (uint32_t)h.w * (uint32_t)cap_y;
        if constexpr (MX) {
          // ---------------------------------------------------------- mixed precision ---
          // Same quads, same lock-step trips; per pair one LDS.U16 (index) and one LDS.128 (record).
          // A quarter-warp = the eight lanes of one row = one shared-memory wavefront when the
          // eight records are distinct modulo 8, which runs of consecutive indices are.
          // Two copies of the loop: most tiles have their five y-rows in consecutive ring slots
          // (record address = one IMAD); the others wrap around the end of the ring (+ ISETP, IADD).
          // G lanes per row, 32/G rows per warp in lock step (a "quad" of the FP64 kernel).  G = 4: eight
          // rows share the ~110 instructions of row set-up and reduction, twice the trips per set-up.
          constexpr int G = kCtLanesMx, R = 32 / G;
          const int lgx = lane & (G - 1), gix = lane / G;
          // the ring offset of this tile, made opaque (a shuffle) so that ptxas keeps it in a register:
          // it otherwise re-derives yslot0 * cap_y inside the pair loop, one IMAD per pair
          const uint32_t c1 = __shfl_sync(0xffffffffu, ybase_s + off0 * 16u, 0);
#if LJ_CT_DIAG
          const int dmode = modev & 15;
#endif
          const int ngroups = nquads;  // groups of R rows, dealt round-robin across tiles like the FP64 quads
          int grp = quad;
          const uint32_t ring_end = ybase_s + ring * 16u, ring_bytes = ring * 16u;
          auto run = [&](auto wrap_tag) {
            constexpr bool WRAP = decltype(wrap_tag)::value;
            auto fetchx = [&](uint32_t L) {
              uint32_t a = L * 16u + c1;
              if (WRAP) { if (a >= ring_end) a -= ring_bytes; }  // past the last slot: back to slot 0
              int4 v;
              asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
              return v;
            };
            for (; grp < ngroups; grp += NCONS) {
              const int r = grp * R + gix;
              const bool valid = r < ns;
              int4 m = make_int4(0, (int)u0, 0, 0);
              if (valid) m = meta[r];
              const int np = m.x;
              const int trips = ((np + 7) >> 3) * (8 / G);  // rows are padded to 8 entries
#if LJ_CT_TMIN_ALL
              const int tmin = __reduce_min_sync(0xffffffffu, trips);
#else
              // rows past the end of the tile (the last quad of three tiles in four) do not shorten the unrolled
              // loop: they walk the tile's first entries against the dummy point (every pair masked)
              const int tmin = __reduce_min_sync(0xffffffffu, valid ? trips : 0x7fffffff);
#endif
              const int tmax = __reduce_max_sync(0xffffffffu, trips);
              const uint16_t* __restrict__ e = lst + ((uint32_t)m.y - u0) * 8u + lgx;
              const int4 me = fetchx(valid ? (uint32_t)(self0 + r) : dummy);
              float fx = 0.f, fy = 0.f, fz = 0.f;
              float nearest = 3.0e38f;  // min |r2 - cl2| over the row's pairs of this lane
#if LJ_CT_ALU_SUB
              const int4 nme = make_int4((int)(0u - (uint32_t)me.x), (int)(0u - (uint32_t)me.y), (int)(0u - (uint32_t)me.z), 0);
#endif
              // r2 from exact differences in counts (modulo 2^32: correct for |d| < 2^31 counts)
              auto dist = [&](const int4 pj, float& dx, float& dy, float& dz) {
#if LJ_CT_ALU_SUB
                // the subtraction as VIADDMNMX (ALU pipe): ptxas otherwise emits IMAD.IADD, and the FMA
                // pipe already carries the 12 FP32 operations of the pair
                dx = (float)__viaddmin_s32(pj.x, nme.x, kbig);
                dy = (float)__viaddmin_s32(pj.y, nme.y, kbig);
                dz = (float)__viaddmin_s32(pj.z, nme.z, kbig);
#else
                dx = (float)(int)((uint32_t)pj.x - (uint32_t)me.x);
                dy = (float)(int)((uint32_t)pj.y - (uint32_t)me.y);
                dz = (float)(int)((uint32_t)pj.z - (uint32_t)me.z);
#endif
                return fmaf(dz, dz, fmaf(dy, dy, dx * dx)) * unit2;
              };
              auto force = [&](float r2) {  // df * unit: the differences stay in counts
                float x;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(r2));
                const float x2 = x * x;
                return (x2 * x2) * fmaf(-c48u, x2 * x, c24u);
              };
              // Hot path: a pair counts iff r2 <= lo_c = cl2 - margin, i.e. it is inside the cutoff
              // whatever the FP32 / fixed-point error.  Pairs in the band around the cutoff are left
              // out here and only REMEMBERED (one FADD + FMNMX, no branch); rows that saw one are
              // revisited below with the exact FP64 test.
              auto pairx = [&](const int4 pj) {
                float dx, dy, dz;
                const float r2 = dist(pj, dx, dy, dz);
                nearest = fminf(nearest, fabsf(r2 - cl2f));
                const float df = r2 <= lo_c ? force(r2) : 0.f;
                fx = fmaf(df, dx, fx);
                fy = fmaf(df, dy, fy);
                fz = fmaf(df, dz, fz);
              };
              int k = 0;
              for (; k + kCtUnrollMx <= tmin; k += kCtUnrollMx) {
                uint32_t en[kCtUnrollMx];
#pragma unroll
                for (int v = 0; v < kCtUnrollMx; v++) en[v] = e[(k + v) * G];
                int4 pj[kCtUnrollMx];
#if LJ_CT_DIAG
                if (dmode == 2) {  // diagnostics: all lanes read one record (no conflicts, no gather)
#pragma unroll
                  for (int v = 0; v < kCtUnrollMx; v++) en[v] = (en[v] & 0u) + dummy;
                }
#endif
#pragma unroll
                for (int v = 0; v < kCtUnrollMx; v++) pj[v] = fetchx(en[v]);
#if LJ_CT_DIAG
                if (dmode == 1) {  // diagnostics: loads only, no pair arithmetic
#pragma unroll
                  for (int v = 0; v < kCtUnrollMx; v++) { fx += __int_as_float(pj[v].x); fy += __int_as_float(pj[v].y); fz += __int_as_float(pj[v].z); }
                  continue;
                }
#endif
#pragma unroll
                for (int v = 0; v < kCtUnrollMx; v++) pairx(pj[v]);
              }
              for (; k < tmax; k += 2) {  // warp-uniform; rows that are already done look at the dummy point
                const uint32_t e0 = k < trips ? (uint32_t)e[k * G] : dummy;
                const uint32_t e1 = k + 1 < trips ? (uint32_t)e[(k + 1) * G] : dummy;
                const int4 p0 = fetchx(e0), p1 = fetchx(e1);
                pairx(p0);
                pairx(p1);
              }
              if (nearest <= P.band) {  // rare (about one row in 300 at rho = 1): the borderline pairs, exactly
                for (int kk = 0; kk < trips; kk++) {
                  const int4 pj = fetchx((uint32_t)e[kk * G]);
                  float dx, dy, dz;
                  const float r2 = dist(pj, dx, dy, dz);
                  if (r2 <= lo_c || !(fabsf(r2 - cl2f) <= P.band)) continue;
                  double xi, yi, zi, xj, yj, zj;
                  load_pos<LAYOUT>(P.q, m.z, P.plane, xi, yi, zi);
                  load_pos<LAYOUT>(P.q, pj.w, P.plane, xj, yj, zj);
                  const double ex = xj - xi, ey = yj - yi, ez = zj - zi;
                  if (fma(ez, ez, fma(ey, ey, ex * ex)) <= P.cl2) {
                    const float df = force(r2);
                    fx = fmaf(df, dx, fx);
                    fy = fmaf(df, dy, fy);
                    fz = fmaf(df, dz, fz);
                  }
                }
              }
              // FP32 per-lane partial sums (about 17 pairs each), FP64 from here on
              const double sx = group_sum<G>((double)fx);
              const double sy = group_sum<G>((double)fy);
              const double sz = group_sum<G>((double)fz);
              if (lgx == 0 && np > 0) red_mom<LAYOUT>(P.p, m.z, P.plane, sx, sy, sz);
              n_quads++;
            }
          };
          if (h.w + kTileYPencils > ry) run(std::true_type{}); else run(std::false_type{});
        } else {
          // ------------------------------------------------------------------ FP64 ---
          const uint32_t zbase_s = smem_u32(zbase);
          // ring offsets of this tile, opaque to ptxas (see the mixed branch)
          const uint32_t c1 = __shfl_sync(0xffffffffu, ybase_s + off0 * 16u, 0);
          const uint32_t c1z = __shfl_sync(0xffffffffu, zbase_s + off0 * 8u, 0);
          const uint32_t ring_end = ybase_s + ring * 16u, ring_xy = ring * 16u, ring_z = ring * 8u;
          auto run = [&](auto wrap_tag) {
            constexpr bool WRAP = decltype(wrap_tag)::value;
            auto fetch = [&](uint32_t L, double& x, double& y, double& z) {
              uint32_t a = L * 16u + c1, az = L * 8u + c1z;
              if (WRAP) { if (a >= ring_end) { a -= ring_xy; az -= ring_z; } }  // past the last slot: back to slot 0
              asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
              asm volatile("ld.shared.f64 %0, [%1];" : "=d"(z) : "r"(az));
            };
            for (; quad < nquads; quad += NCONS) {
              const int r = quad * 4 + gi;
              const bool valid = r < ns;
              int4 m = make_int4(0, (int)u0, 0, 0);
              if (valid) m = meta[r];
              const int np = m.x;
              const int trips = (np + 7) >> 3;
#if LJ_CT_TMIN_ALL
              const int tmin = __reduce_min_sync(0xffffffffu, trips);
#else
              // rows past the end of the tile (the last quad of three tiles in four) do not shorten the unrolled
              // loop: they walk the tile's first entries against the dummy point (every pair masked)
              const int tmin = __reduce_min_sync(0xffffffffu, valid ? trips : 0x7fffffff);
#endif
              const int tmax = __reduce_max_sync(0xffffffffu, trips);
              const uint16_t* __restrict__ e = lst + ((uint32_t)m.y - u0) * 8u + lg;
              double xi, yi, zi;
              fetch(valid ? (uint32_t)(self0 + r) : dummy, xi, yi, zi);
              double fx = 0.0, fy = 0.0, fz = 0.0;
              int k = 0;
              for (; k + kCtUnroll <= tmin; k += kCtUnroll) {
                uint32_t en[kCtUnroll];
#pragma unroll
                for (int v = 0; v < kCtUnroll; v++) en[v] = e[(k + v) * 8];
                double xj[kCtUnroll], yj[kCtUnroll], zj[kCtUnroll];
#pragma unroll
                for (int v = 0; v < kCtUnroll; v++) fetch(en[v], xj[v], yj[v], zj[v]);
#pragma unroll
                for (int v = 0; v < kCtUnroll; v++)
                  lj_pair(xj[v] - xi, yj[v] - yi, zj[v] - zi, P.c24, P.c48, P.cl2_bits, fx, fy, fz);
              }
              for (; k < tmax; k += 2) {  // warp-uniform; rows that are already done look at the dummy point
                const uint32_t e0 = k < trips ? (uint32_t)e[k * 8] : dummy;
                const uint32_t e1 = k + 1 < trips ? (uint32_t)e[(k + 1) * 8] : dummy;
                double x0, y0_, z0, x1, y1_, z1;
                fetch(e0, x0, y0_, z0);
                fetch(e1, x1, y1_, z1);
                lj_pair(x0 - xi, y0_ - yi, z0 - zi, P.c24, P.c48, P.cl2_bits, fx, fy, fz);
                lj_pair(x1 - xi, y1_ - yi, z1 - zi, P.c24, P.c48, P.cl2_bits, fx, fy, fz);
              }
              fx = group_sum<8>(fx);
              fy = group_sum<8>(fy);
              fz = group_sum<8>(fz);
              // RED (no return value): the warp does not wait for p at the end of every quad.  Exactly
              // one add per component, row and step, so the result is deterministic.
              if (lg == 0 && np > 0) red_mom<LAYOUT>(P.p, m.z, P.plane, fx, fy, fz);
              n_quads++;
            }
          };
          if (h.w + kTileYPencils > ry) run(std::true_type{}); else run(std::false_type{});
        }  // FP64
      }
      __syncwarp();
      if (dbgp && had_quad) t_work += clock64() - tw1;
      if (lane == 0) mbar_arrive(&tempty[tslot]);  // this warp is through with the tile
      if (++tslot == rl) { tslot = 0; tphase ^= 1; }
    }
  }
  if (dbgp && lane == 0) {
    long long* d = dbgp + ((size_t)blockIdx.x * NCONS + warp) * 4;
    d[0] = t_wait; d[1] = t_work; d[2] = n_quads; d[3] = clock64() - t_begin;
  }
}

// Ring sizes for a CTA with `budget` bytes of dynamic shared memory.  A tile holds five y-rows; at a unit
// boundary the last tile of the old unit and the first tile of the new one hold ten y-rows between them, so
// fewer than ten y slots drain the pipeline at every boundary, and every further slot is worth ~2 % (measured
// 8 .. 11 slots: 0.342 / 0.334 / 0.324 / 0.318 ms).  The list ring is byte-granular: it must hold two tiles of
// the longest kind (ls_max each), beyond that kCtListDepth AVERAGE tiles are enough and the rest of the
// shared memory goes to y slots.
constexpr int kCtListDepth = LJ_CT_LIST_DEPTH;
static bool ring_sizes(size_t budget, size_t ys, size_t ls_max, size_t ls_avg, int& ry, int& rl, size_t& lring) {
  ry = rl = 0; lring = 0;
  const size_t lmin = 2 * ls_max;
  size_t want = (size_t)kCtListDepth * ls_avg;
  if (want < lmin) want = lmin;
  if (want + kTileMinYSlots * ys > budget) want = lmin;
  if (want + kTileMinYSlots * ys > budget) return false;
  int y = (int)((budget - want) / ys);
  if (y > kCtMaxY) y = kCtMaxY;
  ry = y;
  rl = kCtMaxL;
  lring = (budget - (size_t)ry * ys) & ~(size_t)15;
  return true;
}
// bytes an average tile takes in the list ring (row metadata + list units)
static size_t avg_tile_bytes(const lj_ctx* ctx) {
  const lj_tile_geom& g = ctx->tl_g;
  const size_t tiles = g.ntiles > 0 ? (size_t)g.ntiles : 1;
  size_t avg = 16 * ((size_t)g.total_units + (size_t)ctx->tl_pn) / tiles;
  const size_t ls = lj_celltile_lslot_bytes(g);
  if (avg < ls / 4) avg = ls / 4;  // (empty tiles in the count pull the mean down)
  return (avg + 15) & ~(size_t)15;
}

template <int LAYOUT, bool MX, int NCONS, int NB>
int launch_celltile(lj_ctx* ctx, const lj_force_args* a, double c24, double c48, long long cl2_bits,
                    cudaStream_t st, int ry, int rl, size_t lring, int part) {
  const lj_tile_geom& g = ctx->tl_g;
  const size_t ys = (size_t)lj_celltile_cap_y(g) * (MX ? 16 : 24), ls = lj_celltile_lslot_bytes(g);
  {  // diagnostics: cap the ring sizes (LJ_TILE_LRING in KB, not below two of the longest tiles)
    const int ry_env = lj_diag_int("LJ_TILE_RY"), rl_env = lj_diag_int("LJ_TILE_RL"), lr_env = lj_diag_int("LJ_TILE_LRING");
    if (ry_env >= kTileMinYSlots && ry_env < ry) ry = ry_env;
    if (rl_env >= kTileMinLSlots && rl_env < rl) rl = rl_env;
    if (lr_env > 0 && (size_t)lr_env * 1024 < lring && (size_t)lr_env * 1024 >= 2 * ls) lring = (size_t)lr_env * 1024;
  }
  const int seg_env = lj_diag_int("LJ_TILE_SEG");
  // columns without a single list entry (the ghost layers of a decomposed run) are skipped
  const bool all_cols = g.ncols_active <= 0 || g.ncols_active >= g.ntx * g.nz;
  const int ncols = all_cols ? g.ntx * g.nz : g.ncols_active;
  int nseg = (16 * NB * ctx->sm_count + ncols - 1) / ncols;  // >= 16 units per CTA: the dynamic deal ends evenly
  if (nseg > g.ny / 6) nseg = g.ny / 6;                 // but every unit re-stages four y-rows: keep them long
  if (nseg < 1) nseg = 1;
  int seg_len = seg_env > 0 ? seg_env : (g.ny + nseg - 1) / nseg;
  if (seg_len > g.ny) seg_len = g.ny;
  if (seg_len > kCtMaxSeg) seg_len = kCtMaxSeg;
  nseg = (g.ny + seg_len - 1) / seg_len;
  // Units are dealt dynamically in the order (segment, column).  With equal segments the kernel ends when the
  // CTA that drew the last 11-tile unit is through: CTAs finished between 571k and 606k cycles (6 % apart).
  // GRADED segments -- long ones first, each 45 % of what is left, the last ones 4-5 tiles (ny = 63: 29, 16, 9, 5, 4) -- end the
  // deal within a few tiles of each other and re-stage fewer y-rows on the way (every unit stages four extra).
  short seg_y0[kCtMaxUnitsPerCol + 1];
  bool graded = false;
  if (seg_env <= 0 && kCtGraded && nseg >= 3 && nseg <= 8) {
    int lens[kCtMaxUnitsPerCol], n = 0, rem = g.ny;
    while (rem > 0 && n < kCtMaxUnitsPerCol) {
      int len = (rem * LJ_CT_GRADE_PCT + 99) / 100;
      if (len < LJ_CT_GRADE_MIN) len = LJ_CT_GRADE_MIN;
      if (len > kCtMaxSeg) len = kCtMaxSeg;
      if (rem - len < 3 && rem <= kCtMaxSeg) len = rem;
      if (len > rem) len = rem;
      lens[n++] = len;
      rem -= len;
    }
    if (rem > 0) n = 0;  // (does not happen: ny <= 32 * 32) fall back to equal segments
    if (n > 0) {
      std::sort(lens, lens + n, [](int x, int y) { return x > y; });
      nseg = n; seg_len = lens[0];
      seg_y0[0] = 0;
      for (int k = 0; k < n; k++) seg_y0[k + 1] = (short)(seg_y0[k] + lens[k]);
      graded = true;
    }
  }
  LJ_REQUIRE(ctx, nseg <= kCtMaxUnitsPerCol, "cell-tile force: too many segments per column");
  if (!graded)
    for (int k = 0; k <= nseg; k++) seg_y0[k] = (short)(k * seg_len < g.ny ? k * seg_len : g.ny);

  ct_params P;
  P.qs = MX ? reinterpret_cast<const unsigned char*>(ctx->tl_qfx) : reinterpret_cast<const unsigned char*>(ctx->tl_qs);
  P.qz = reinterpret_cast<const unsigned char*>(ctx->tl_qz);
  P.p = a->p; P.plane = a->plane_stride;
  P.c24 = c24; P.c48 = c48; P.cl2_bits = cl2_bits;
  P.q = a->q; P.grid = ctx->grid; P.cl2 = a->cl2;
  {
    const lj_fx_frame f = lj_fx_frame_for(a->cl2);
    P.fx_scale = f.scale;
    P.c24u = (float)(c24 * f.unit); P.c48u = (float)(c48 * f.unit);
    P.unit2 = (float)(f.unit * f.unit);  // a power of two: exact
    P.cl2f = (float)a->cl2; P.lo_c = P.cl2f - f.margin; P.band = 1.01f * f.margin;
  }
  P.ytab = ctx->tl_tab; P.ttab = ctx->tl_ttab; P.meta = ctx->tl_meta; P.list = ctx->tl_list;
  P.ntx = g.ntx; P.ny = g.ny; P.ncols = ncols; P.seg_len = seg_len; P.nseg = nseg;
  for (int k = 0; k <= nseg; k++) P.seg_y0[k] = seg_y0[k];
  P.cols = all_cols ? nullptr : ctx->tl_cols;
  P.ncols_dev = nullptr;
  if (part != 0) {  // the permute kernel compacted the selected columns; their number stays on the device
    P.cols = ctx->tl_cols_sel;
    P.ncols_dev = &ctx->tl_geom->pad2;
  }
  P.cap_y = lj_celltile_cap_y(g); P.cap_units = g.max_units; P.cap_rows = g.max_rows;
  P.ry = ry; P.rl = rl; P.lring_bytes = (int)lring;
  P.mode = lj_diag_int("LJ_TILE_MODE");
  P.unit_counter = &ctx->tl_geom->pad;
  P.dbg = nullptr;
#if LJ_DIAG
  if (lj_diag_set("LJ_TILE_DBG")) {
    if (!ctx->diag_buf) cudaMalloc(&ctx->diag_buf, sizeof(long long) * 4 * 32 * 1024);
    P.dbg = ctx->diag_buf;
  }
#endif
  const size_t smem = (size_t)ry * ys + lring;
  auto kern = lj_celltile_force<LAYOUT, MX, NCONS, NB>;
  LJ_FUNC_SMEM(ctx, kern, smem);
  const int nunits = ncols * nseg;  // part launches: an upper bound, CTAs without a unit leave at once
  const int grid = nunits < NB * ctx->sm_count ? nunits : NB * ctx->sm_count;
  if (lj_diag_set("LJ_TILE_DEBUG"))
    fprintf(stderr, "[lj] cell-tile force: %d consumer warps, %d units (%d columns x %d segments of %d), "
            "y ring %d x %zu B, list ring %zu B for %d tiles in flight (longest tile %zu B), smem %zu B\n", NCONS, nunits,
            ncols, nseg, seg_len, ry, ys, lring, rl, ls, smem);
  cudaEvent_t kt0 = lj_kernel_timing_event(ctx, st, true);  // lj_kernel_timing: this kernel alone, live
  if (kt0) cudaEventRecord(kt0, st);
  kern<<<(unsigned)grid, (NCONS + 2) * 32, smem, st>>>(P);
  if (kt0) cudaEventRecord(lj_kernel_timing_event(ctx, st, false), st);
  LJ_LAUNCHED(ctx);
#if LJ_DIAG
  if (P.dbg) {  // diagnostics only: synchronises
    cudaStreamSynchronize(st);
    if (ctx->diag_dumps++ == 3) {
      std::vector<long long> h((size_t)4 * (NCONS + 2) * grid);
      cudaMemcpy(h.data(), P.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      double w = 0, k = 0, q = 0, t = 0, tmax = 0, tmin = 1e30, qmax = 0, qmin = 1e30;
      double wlo = 0, whi = 0, klo = 0, khi = 0;  // per CTA: least / most waiting warp, least / most working warp
      for (int b = 0; b < grid; b++) {
        double qb = 0, tb = 0, w0 = 1e30, w1 = 0, k0 = 1e30, k1 = 0;
        for (int c = 0; c < NCONS; c++) {
          const long long* d = &h[((size_t)b * NCONS + c) * 4];
          w += d[0]; k += d[1]; q += d[2]; t += d[3]; qb += d[2]; if (d[3] > tb) tb = d[3];
          if (d[0] < w0) w0 = d[0]; if (d[0] > w1) w1 = d[0]; if (d[1] < k0) k0 = d[1]; if (d[1] > k1) k1 = d[1];
        }
        wlo += w0; whi += w1; klo += k0; khi += k1;
        if (tb > tmax) tmax = tb; if (tb < tmin) tmin = tb; if (qb > qmax) qmax = qb; if (qb < qmin) qmin = qb;
      }
      for (int pw = 0; pw < 2; pw++) {
        double pi = 0, pt = 0, ptl = 0;
        for (int b = 0; b < grid; b++) {
          const long long* d = &h[((size_t)grid * NCONS + 2 * b + pw) * 4];
          pi += d[0]; ptl += d[2]; pt += d[3];
        }
        fprintf(stderr, "[lj] celltile dbg: producer warp %c avg idle %.0f of %.0f cycles, %.1f tiles per CTA -> busy %.0f cycles per tile\n",
                pw ? 'L' : 'Y', pi / grid, pt / grid, ptl / grid, (pt - pi) / ptl);
      }
      const double n = (double)grid * NCONS;
      fprintf(stderr, "[lj] celltile dbg: per warp avg wait %.0f, work %.0f, total %.0f cycles, quads %.1f (%.0f cycles/quad); "
              "CTA total min %.0f max %.0f, quads per CTA min %.0f max %.0f\n", w / n, k / n, t / n, q / n, k / q, tmin, tmax, qmin, qmax);
      fprintf(stderr, "[lj] celltile dbg: within a CTA (avg over CTAs): wait of the least / most waiting warp %.0f / %.0f, "
              "work of the least / most working warp %.0f / %.0f\n", wlo / grid, whi / grid, klo / grid, khi / grid);
    }
  }
#endif
  return LJ_OK;
}

// LJ_CT_NCONS (FP64) / LJ_CT_NCONS_MX (mixed) consumer warps + 2 producer warps in one CTA per SM; other layouts in
// diagnostic builds only
template <int LAYOUT, bool MX>
int dispatch_celltile(lj_ctx* ctx, const lj_force_args* a, double c24, double c48, long long cl2_bits,
                      cudaStream_t st, int part) {
  const lj_tile_geom& g = ctx->tl_g;
  const size_t ys = (size_t)lj_celltile_cap_y(g) * (MX ? 16 : 24), ls = lj_celltile_lslot_bytes(g);
  const size_t la = avg_tile_bytes(ctx);
  int ry = 0, rl = 0;
  size_t lring = 0;
#if LJ_DIAG
  const int nc = lj_diag_int("LJ_TILE_CONSUMERS");
  if (nc == 8) {  // two CTAs per SM (measured slower: two pipelines, twice the producers)
    const size_t half = (size_t)(227 * 1024) / 2 - 9 * 1024;
    LJ_REQUIRE(ctx, ring_sizes(half, ys, ls, la, ry, rl, lring), "lj_force_step: cell-tile geometry does not fit in shared memory");
    return launch_celltile<LAYOUT, MX, 8, 2>(ctx, a, c24, c48, cl2_bits, st, ry, rl, lring, part);
  }
  if (nc == 24 && ring_sizes(kTileSmemBudget, ys, ls, la, ry, rl, lring))
    return launch_celltile<LAYOUT, MX, 24, 1>(ctx, a, c24, c48, cl2_bits, st, ry, rl, lring, part);
  if (nc == 20 && ring_sizes(kTileSmemBudget, ys, ls, la, ry, rl, lring))
    return launch_celltile<LAYOUT, MX, 20, 1>(ctx, a, c24, c48, cl2_bits, st, ry, rl, lring, part);
#endif
  LJ_REQUIRE(ctx, ring_sizes(kTileSmemBudget, ys, ls, la, ry, rl, lring), "lj_force_step: cell-tile geometry does not fit in shared memory");
  return launch_celltile<LAYOUT, MX, MX ? LJ_CT_NCONS_MX : LJ_CT_NCONS, 1>(ctx, a, c24, c48, cl2_bits, st, ry, rl, lring, part);
}

}  // namespace

// AUTO takes the cell-tile kernel only where it wins: the persistent CTAs need a few dozen tiles each
// to amortise their pipeline fill (measured: slower than the per-row kernel at N = 23k, faster at 1M)
bool lj_celltile_worthwhile(const lj_ctx* ctx) { return ctx->tl_g.ntiles >= 32 * ctx->sm_count; }

// true when the mirror describes exactly the list arrays and the row range of this call
bool lj_celltile_usable(const lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1) {
  if (!ctx->tl_valid || a->list_layout != LJ_LIST_CSR) return false;
  // pointer identity alone cannot tell a list that was overwritten in place: the caller vouches with the token
  if (a->mirror_token != ctx->tl_token) return false;
  if (ctx->tl_outside > 0 && a->precision != LJ_PREC_FP64) return false;  // rows outside the mirror: FP64 per-row completion only
  if (a->precision != LJ_PREC_FP64 && a->precision != LJ_PREC_MIXED) return false;
  if (a->layout != LJ_AOS_D3 && a->layout != LJ_AOS_D4 && a->layout != LJ_SOA_D) return false;
  if (a->list != ctx->tl_id_list || a->number_of_partners != ctx->tl_id_nop ||
      a->pointer != ctx->tl_id_ptr || a->pn != ctx->tl_pn)
    return false;
  return r0 == ctx->tl_r0 && r1 == ctx->tl_r1;
}

int lj_force_celltile_launch(lj_ctx* ctx, const lj_force_args* a, double c24, double c48,
                             long long cl2_bits, cudaStream_t st, int part) {
  int rc = lj_celltile_permute(ctx, a, st, part);
  if (rc) return rc;
  const bool mx = a->precision == LJ_PREC_MIXED;
  switch (a->layout) {
    case LJ_AOS_D4:
      return mx ? dispatch_celltile<LJ_AOS_D4, true>(ctx, a, c24, c48, cl2_bits, st, part)
                : dispatch_celltile<LJ_AOS_D4, false>(ctx, a, c24, c48, cl2_bits, st, part);
    case LJ_AOS_D3:
      return mx ? dispatch_celltile<LJ_AOS_D3, true>(ctx, a, c24, c48, cl2_bits, st, part)
                : dispatch_celltile<LJ_AOS_D3, false>(ctx, a, c24, c48, cl2_bits, st, part);
    case LJ_SOA_D:
      return mx ? dispatch_celltile<LJ_SOA_D, true>(ctx, a, c24, c48, cl2_bits, st, part)
                : dispatch_celltile<LJ_SOA_D, false>(ctx, a, c24, c48, cl2_bits, st, part);
  }
  return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_force_step", "layout");
}
