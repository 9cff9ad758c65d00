// lj_nlist.cu -- on-GPU Verlet neighbour-list build (cell binning + stencil search + scan).
//
// Replaces makepair()/register_pair() (cuda/force_cuda.cu:102-163), an O(N^2) host loop that
// materialises (i,j) pair arrays, with an O(N) device pipeline that never leaves the GPU:
//
//   bbox -> grid setup (device) -> cell id + arrival slot (atomics) -> scan(cell counts)
//        -> scatter -> deterministic in-cell ordering -> COUNT pass (27-cell stencil)
//        -> scan(number_of_partners) = pointer[] (64-bit inside) -> FILL pass
//
// Output contract = the reference's: number_of_partners[i], pointer[] (exclusive scan, pn
// entries), sorted_list in the caller's numbering.  Membership is decided by the exact FP64
// expression r2 = fma(dz,dz,fma(dy,dy,dx*dx)) < search^2 (same chain as oracle/lj_oracle.c);
// an FP32 test on origin-shifted coordinates only pre-classifies candidates that are farther
// than a rigorous error margin from the threshold.
#include <cstdlib>

#include "lj_celltile.cuh"
#include "lj_common.cuh"

#ifndef LJ_TILE_ROWS_WIDE
#define LJ_TILE_ROWS_WIDE 72
#endif
#ifndef LJ_TILE_ROWS_STD
#define LJ_TILE_ROWS_STD 56
#endif

namespace {

// ------------------------------------------------------------------ small utilities ---
__global__ void k_bbox_init(unsigned long long* bb, lj_list_totals* tot) {
  if (threadIdx.x < 3) bb[threadIdx.x] = ~0ull;      // mins
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0ull;  // maxs
  if (threadIdx.x == 0 && tot) { tot->total = 0; tot->max_np = 0; tot->overflow = 0; tot->cl_total = 0; }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_bbox(const void* __restrict__ q, int64_t pn, int64_t plane, unsigned long long* bb) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  // four independent loads in flight per thread: one load per trip left the kernel latency-bound
  // (24 us for 32 MB at N = 1M)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < pn; i += 4 * stride) {
    double v[4][3];
#pragma unroll
    for (int u = 0; u < 4; u++) load_pos<LAYOUT>(q, i + u * stride, plane, v[u][0], v[u][1], v[u][2]);
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int d = 0; d < 3; d++) { lo[d] = fmin(lo[d], v[u][d]); hi[d] = fmax(hi[d], v[u][d]); }
  }
  for (; i < pn; i += stride) {
    double v[3];
    load_pos<LAYOUT>(q, i, plane, v[0], v[1], v[2]);
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[d] = fmin(lo[d], v[d]); hi[d] = fmax(hi[d], v[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; d++)
    for (int m = 16; m >= 1; m >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], m));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], m));
    }
  __shared__ double wlo[8][3], whi[8][3];  // one set of six atomics per block, not per warp
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { wlo[warp][d] = lo[d]; whi[warp][d] = hi[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    double l = wlo[0][d], h = whi[0][d];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { l = fmin(l, wlo[w][d]); h = fmax(h, whi[w][d]); }
    atomicMin(bb + d, enc_ordered(l));
    atomicMax(bb + 3 + d, enc_ordered(h));
  }
}

struct grid_ext {  // lj_grid_params + the FP32 pre-filter constants, lives right behind it
  lj_grid_params g;
  int ncell1;      // ncell + 1: the scanned histogram carries a sentinel (cell_start[ncell] = pn)
  float margin;    // |r2_f32 - r2_f64| bound for stencil candidates
  float sl2f;      // search^2
  float edge;      // cell edge (>= search/2)
  float inv_edge;
  float pad;       // bound on the float error of a shifted coordinate / cell boundary
};

// Cells of edge >= search/2 and a 5x5x5 stencil: 4.6x the sphere volume instead of the 7.7x of
// search-sized cells with 27 neighbours; the x-chord trimming in k_search brings it to ~2.5x.
__global__ void k_grid_setup(const unsigned long long* bb, double search_len, int64_t cap_cells,
                             grid_ext* out) {
  double lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = dec_ordered(bb[d]); hi[d] = dec_ordered(bb[3 + d]); }
  double edge = 0.5 * search_len * (1.0 + 1e-9);  // strictly larger: no neighbour 3 cells away
  int n[3];
  for (;;) {
    double cells = 1.0;
    for (int d = 0; d < 3; d++) {
      double c = floor((hi[d] - lo[d]) / edge) + 1.0;
      if (c > 2.0e9) c = 2.0e9;
      n[d] = (int)c;
      cells *= c;
    }
    if (cells + 1.0 <= (double)cap_cells) break;
    edge *= 1.26;  // sparse cloud: coarser cells stay correct for the +-2 stencil
  }
  out->g.ox = lo[0]; out->g.oy = lo[1]; out->g.oz = lo[2];
  out->g.inv_cell = 1.0 / edge;
  out->g.nx = n[0]; out->g.ny = n[1]; out->g.nz = n[2];
  out->g.ncell = n[0] * n[1] * n[2];
  out->ncell1 = out->g.ncell + 1;
  // FP32 pre-filter error budget.  E = largest extent; a shifted coordinate rounds to float
  // with error <= 2^-24 E, a float difference of two of them adds <= 2^-24 |d|, |d| <= 3 edge
  // for stencil candidates.  r2 error <= sum_c (2|d| delta + delta^2) + rounding of the
  // three multiply-adds.  Doubled for safety.
  double E = fmax(hi[0] - lo[0], fmax(hi[1] - lo[1], hi[2] - lo[2]));
  const double u = 5.9604644775390625e-8;  // 2^-24
  double delta = 2.0 * u * E + u * 3.0 * edge;
  double m = 3.0 * (6.0 * edge * delta + delta * delta) + 4.0 * u * 27.0 * edge * edge;
  out->margin = (float)(2.0 * m);
  out->sl2f = (float)(search_len * search_len);
  out->edge = (float)edge;
  out->inv_edge = (float)(1.0 / edge);
  out->pad = (float)(8.0 * u * (E + edge) + 1e-30);
}

__device__ __forceinline__ int cell_coord(double v, double o, double inv, int n) {
  int c = (int)floor((v - o) * inv);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_cell_assign(const void* __restrict__ q, int64_t pn, int64_t plane, const grid_ext* __restrict__ ge,
              int32_t* __restrict__ cell_of, int32_t* __restrict__ cell_slot,
              uint32_t* __restrict__ cell_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  const lj_grid_params g = ge->g;
  double x, y, z;
  load_pos<LAYOUT>(q, i, plane, x, y, z);
  const int cx = cell_coord(x, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(y, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(z, g.oz, g.inv_cell, g.nz);
  const int c = (cz * g.ny + cy) * g.nx + cx;
  cell_of[i] = c;
  cell_slot[i] = (int32_t)atomicAdd(cell_count + c, 1u);
}

__global__ void __launch_bounds__(256)
k_cell_scatter(int64_t pn, const int32_t* __restrict__ cell_of, const int32_t* __restrict__ cell_slot,
               const uint32_t* __restrict__ cell_start, int32_t* __restrict__ sorted_tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  sorted_tmp[cell_start[cell_of[i]] + (uint32_t)cell_slot[i]] = (int32_t)i;
}

// Arrival order inside a cell depends on atomic timing; re-rank by original index so the
// whole build is deterministic: rank = #members of my cell with a smaller index.
template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_cell_order(const void* __restrict__ q, int64_t pn, int64_t plane, const grid_ext* __restrict__ ge,
             const int32_t* __restrict__ cell_of, const uint32_t* __restrict__ cell_start,
             const uint32_t* __restrict__ cell_count, const int32_t* __restrict__ sorted_tmp,
             double4* __restrict__ sorted_pos, float4* __restrict__ sorted_pos32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  const int c = cell_of[i];
  const uint32_t b = cell_start[c], n = cell_count[c];
  uint32_t rank = 0;
  for (uint32_t m = 0; m < n; m++) rank += (sorted_tmp[b + m] < (int32_t)i) ? 1u : 0u;
  double x, y, z;
  load_pos<LAYOUT>(q, i, plane, x, y, z);
  sorted_pos[b + rank] = make_double4(x, y, z, __longlong_as_double((long long)i));
  const lj_grid_params g = ge->g;
  sorted_pos32[b + rank] = make_float4((float)(x - g.ox), (float)(y - g.oy), (float)(z - g.oz),
                                       __int_as_float((int)i));
}

// ------------------------------------------------------------------ prefix scans ------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v,
                                                                   unsigned long long* total) {
  __shared__ unsigned long long warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, m);
    if (lane >= m) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned long long s = lane < kScanThreads / 32 ? warp_sums[lane] : 0ull;
#pragma unroll
    for (int m = 1; m < kScanThreads / 32; m <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, s, m);
      if (lane >= m) s += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const unsigned long long base = w ? warp_sums[w - 1] : 0ull;
  *total = warp_sums[kScanThreads / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

// n_dev (optional) overrides n with a device-side length (the cell count is only known there)
__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const uint32_t* __restrict__ in, int64_t n, const int* __restrict__ n_dev,
              unsigned long long* __restrict__ tile_sums) {
  if (n_dev) n = *n_dev;
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  unsigned long long s = 0;
  if (base < n) {
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      const int64_t idx = base + (int64_t)k * kScanThreads + threadIdx.x;
      if (idx < n) s += in[idx];
    }
  }
  unsigned long long tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_spine(unsigned long long* __restrict__ tile_sums, int64_t ntiles,
             unsigned long long* __restrict__ total_out) {
  unsigned long long carry = 0;
  for (int64_t b = 0; b < ntiles; b += kScanThreads) {
    const int64_t idx = b + threadIdx.x;
    const unsigned long long v = idx < ntiles ? tile_sums[idx] : 0ull;
    unsigned long long tot;
    const unsigned long long ex = block_exclusive_scan(v, &tot);
    if (idx < ntiles) tile_sums[idx] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// OUT = uint32_t (cell starts / int32 pointer[]) or long long (int64 pointer[])
template <typename OUT>
__global__ void __launch_bounds__(kScanThreads)
k_scan_down(const uint32_t* __restrict__ in, int64_t n, const int* __restrict__ n_dev,
            const unsigned long long* __restrict__ tile_sums, OUT* __restrict__ out) {
  if (n_dev) n = *n_dev;
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  if (base >= n) return;
  // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems)
  uint32_t v[kScanItems];
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int64_t idx = base + (int64_t)threadIdx.x * kScanItems + k;
    v[k] = idx < n ? in[idx] : 0u;
    s += v[k];
  }
  unsigned long long tot;
  unsigned long long run = tile_sums[blockIdx.x] + block_exclusive_scan(s, &tot);
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int64_t idx = base + (int64_t)threadIdx.x * kScanItems + k;
    if (idx < n) out[idx] = (OUT)run;
    run += v[k];
  }
}

// ------------------------------------------------------------------ stencil search ----
// kSearchLanes lanes per particle (visited in cell order), 4 particles per warp.  For each of
// the 25 (dz,dy) rows of the +-2 stencil the x-adjacent cells are ONE contiguous range of
// sorted_pos32; the range is trimmed to the chord of the search sphere at that (dy,dz) before
// the group streams it with 16 B loads.  Classification per candidate: FP32 r2 on origin-shifted
// coordinates; only candidates within the representation error of the threshold (or of zero:
// the particle itself / coincident particles) take the exact FP64 path.  Hits are compacted
// with a group ballot.  FILL=false counts, FILL=true writes the row.
constexpr int kSearchLanes = 8;
#ifndef LJ_SEARCH_MIN_BLOCKS
#define LJ_SEARCH_MIN_BLOCKS 4  // 64 registers, no spills: 2.09 ms vs 2.21 ms per build at N=1M
#endif

template <bool FILL, bool PTR64>
__global__ void __launch_bounds__(256)
k_search(int64_t pn, const grid_ext* __restrict__ ge, const int32_t* __restrict__ cell_of,
         const uint32_t* __restrict__ cell_start,
         const double4* __restrict__ sorted_pos, const float4* __restrict__ sorted_pos32,
         double sl2, int half, int64_t row_begin, int64_t row_end,
         int32_t* __restrict__ nop, const void* __restrict__ pointer, int32_t* __restrict__ list,
         int64_t capacity, lj_list_totals* __restrict__ tot) {
  constexpr int GL = kSearchLanes;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GL;  // slot in cell order
  const int lane = threadIdx.x & 31;
  const int lg = lane % GL;
  const unsigned gbits = ((1u << GL) - 1u) << (lane - lg);
  const unsigned lt_mask = (1u << lane) - 1u;
  // The four groups of a warp stay CONVERGED: every loop below has a warp-uniform trip count
  // and inactive / finished groups ride along predicated off.  (Letting groups diverge made
  // each instruction issue once per group: 4x the instruction count, ncu round 1.)
  bool active = s < pn;
  float4 me32 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) me32 = sorted_pos32[s];
  const int i = __float_as_int(me32.w);
  active = active && i >= row_begin && i < row_end;
  const lj_grid_params g = ge->g;
  const float margin = ge->margin, sl2f = ge->sl2f, edge = ge->edge, pad = ge->pad;
  const float lo_f = sl2f - margin, hi_f = sl2f + margin;
  const int c = active ? cell_of[i] : 0;
  const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);

  int64_t base = 0;
  if (FILL && active) {
    base = row_offset<PTR64>(pointer, i);
    if (base + nop[i] > capacity) {
      if (lg == 0) atomicOr(&tot->overflow, 1);
      active = false;
    }
  }
  // squared gap between me and the slab of cells at offset o = -2..2 along each axis (0 for my
  // own slab, +inf outside the grid), shrunk by the float error bound
  const float kInf = __int_as_float(0x7f800000);
  float gx2[5], gy2[5], gz2[5];
#pragma unroll
  for (int o = 0; o < 5; o++) {
    const int x = cx + o - 2, y = cy + o - 2, z = cz + o - 2;
    const float ax = fmaxf(fmaxf(x * edge - me32.x, me32.x - (x + 1) * edge) - pad, 0.f);
    const float ay = fmaxf(fmaxf(y * edge - me32.y, me32.y - (y + 1) * edge) - pad, 0.f);
    const float az = fmaxf(fmaxf(z * edge - me32.z, me32.z - (z + 1) * edge) - pad, 0.f);
    gx2[o] = (active && x >= 0 && x < g.nx) ? ax * ax : kInf;
    gy2[o] = (active && y >= 0 && y < g.ny) ? ay * ay : kInf;
    gz2[o] = (active && z >= 0 && z < g.nz) ? az * az : kInf;
  }
  int count = 0;
  double4 me = make_double4(0, 0, 0, 0);
  bool have_me = false;

#pragma unroll
  for (int dz = 0; dz < 5; dz++) {
#pragma unroll
    for (int dy = 0; dy < 5; dy++) {
      // chord of the search sphere in this (dy,dz) row of cells: x-cells whose gap fits
      const float rem = hi_f - gz2[dz] - gy2[dy];
      const bool run = gx2[2] < rem;  // my own x-slab (gap 0) is in: rem > 0
      if (!__any_sync(0xffffffffu, run)) continue;
      uint32_t m = 0, m_end = 0;
      if (run) {
        const int xa = cx - ((gx2[0] < rem) ? 2 : (gx2[1] < rem) ? 1 : 0);
        const int xb = cx + ((gx2[4] < rem) ? 2 : (gx2[3] < rem) ? 1 : 0);
        const int rowc = ((cz + dz - 2) * g.ny + (cy + dy - 2)) * g.nx;
        m = cell_start[rowc + xa] + lg;
        m_end = cell_start[rowc + xb + 1];  // sentinel-terminated scan
      }
      while (__any_sync(0xffffffffu, m < m_end)) {
        bool hit = false;
        int j = 0;
        if (m < m_end) {
          const float4 c32 = sorted_pos32[m];
          j = __float_as_int(c32.w);
          const float dx = me32.x - c32.x, dy_ = me32.y - c32.y, dz_ = me32.z - c32.z;
          const float r2f = fmaf(dz_, dz_, fmaf(dy_, dy_, dx * dx));
          if (r2f < hi_f) {
            if (r2f < lo_f && r2f > margin) {
              hit = true;
            } else {  // near the threshold, or near zero (myself / coincident): exact FP64
              if (!have_me) { me = sorted_pos[s]; have_me = true; }
              const double4 cj = sorted_pos[m];
              const double ddx = me.x - cj.x, ddy = me.y - cj.y, ddz = me.z - cj.z;
              hit = (j != i) && (fma(ddz, ddz, fma(ddy, ddy, ddx * ddx)) < sl2);
            }
            if (half) hit = hit && (j > i);
          }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit) & gbits;
        if (FILL && hit) list[base + count + __popc(ballot & lt_mask)] = j;
        count += __popc(ballot);
        m += GL;
      }
    }
  }
  if (!FILL && active && lg == 0) {
    nop[i] = count;
    if (count > *(volatile int*)&tot->max_np) atomicMax(&tot->max_np, count);
  }
}

// ------------------------------------------------------------------ cluster search -----
// Same search, organised around CLUSTERS of four consecutive particles (rows r0+4c .. r0+4c+3):
// one candidate stream per cluster, each candidate tested against the four cluster members.
// Emits, from one pass over the candidates,
//   * the reference's CSR arrays (number_of_partners / sorted_list rows of the four members),
//   * the library-owned CLUSTER PAIR LIST: the union of the four rows, one packed entry per
//     candidate  (member mask << 28) | j,  which lets the cluster force kernel gather q[j] once
//     and use it for up to four i-particles from registers.
// Eight lanes per cluster, four clusters per warp, warp-uniform control flow as in k_search.
// Candidate ranges are derived from the cluster's bounding box, so any particle order is
// handled correctly; it is efficient when consecutive particles are spatially close (lattice
// order), which is what the cluster force kernel needs anyway.
template <bool FILL, bool PTR64, int LAYOUT>
__global__ void __launch_bounds__(256, LJ_SEARCH_MIN_BLOCKS)
k_search_cluster(const void* __restrict__ q, int64_t plane, int64_t pn, const grid_ext* __restrict__ ge,
                 const uint32_t* __restrict__ cell_start, const double4* __restrict__ sorted_pos,
                 const float4* __restrict__ sorted_pos32, double sl2, int half, int64_t row_begin,
                 int64_t row_end, int32_t* __restrict__ nop, const void* __restrict__ pointer,
                 int32_t* __restrict__ list, int64_t capacity, uint32_t* __restrict__ cl_cnt,
                 const long long* __restrict__ cl_ptr, uint32_t* __restrict__ cl_list,
                 int64_t cl_cap, lj_list_totals* __restrict__ tot) {
  constexpr int GL = kSearchLanes;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GL;  // cluster index
  const int lane = threadIdx.x & 31;
  const int lg = lane % GL;
  const unsigned gbits = ((1u << GL) - 1u) << (lane - lg);
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t i0 = row_begin + 4 * c;
  const int nrows = (int)max((int64_t)0, min((int64_t)4, row_end - i0));  // members of this cluster
  bool active = nrows > 0;
  const lj_grid_params g = ge->g;
  const float margin = ge->margin, sl2f = ge->sl2f, edge = ge->edge, inv_edge = ge->inv_edge,
              pad = ge->pad;
  const float lo_f = sl2f - margin, hi_f = sl2f + margin;
  // sure-hit band margin < r2 < lo_f as |r2 - mid_s| < hw_s, shrunk so that the rounding of the
  // subtraction cannot admit a value outside the band
  const float mid_s = 0.5f * (lo_f + margin);
  const float hw_s = 0.5f * (lo_f - margin) * (1.0f - 4.0e-6f);
  const float search_f = sqrtf(hi_f) + pad;

  // member positions, origin-shifted floats (same expression as k_cell_order: bit-identical)
  float px[4], py[4], pz[4];
  float bx0 = 3.0e38f, by0 = 3.0e38f, bz0 = 3.0e38f, bx1 = -3.0e38f, by1 = -3.0e38f, bz1 = -3.0e38f;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    px[r] = py[r] = pz[r] = 3.0e18f;  // absent member: far away from everything
    if (r < nrows) {
      double x, y, z;
      load_pos<LAYOUT>(q, i0 + r, plane, x, y, z);
      px[r] = (float)(x - g.ox); py[r] = (float)(y - g.oy); pz[r] = (float)(z - g.oz);
      bx0 = fminf(bx0, px[r]); bx1 = fmaxf(bx1, px[r]);
      by0 = fminf(by0, py[r]); by1 = fmaxf(by1, py[r]);
      bz0 = fminf(bz0, pz[r]); bz1 = fmaxf(bz1, pz[r]);
    }
  }
  int64_t base[4] = {0, 0, 0, 0};
  long long cbase = 0;
  bool emit_cl = cl_cnt != nullptr;
  if (FILL && active) {
#pragma unroll
    for (int r = 0; r < 4; r++)
      if (r < nrows) {
        base[r] = row_offset<PTR64>(pointer, i0 + r);
        if (base[r] + nop[i0 + r] > capacity) active = false;
      }
    if (!active && lg == 0) atomicOr(&tot->overflow, 1);
    if (emit_cl) {
      cbase = cl_ptr[c];
      if (cl_ptr[c + 1] > cl_cap) {
        emit_cl = false;
        if (lg == 0) atomicOr(&tot->overflow, 4);
      }
    }
  }
  // A cluster whose members are far apart (particle order without spatial coherence) would make
  // the bounding-box candidate set explode; such LOOSE clusters are searched one member per pass
  // (the member's own position as the box, only its bit live).  Union entries may then repeat a j
  // for different members, which the cluster force kernel does not mind.
  const bool loose = active && (bx1 - bx0 > 2.f * edge || by1 - by0 > 2.f * edge || bz1 - bz0 > 2.f * edge);
  const int npass = loose ? nrows : 1;
  int pass = 0;
  unsigned live = loose ? 1u : 0xfu;
  // rows/slabs of cells that can hold a neighbour of the (cluster's or member's) bounding box
  int y = 0, z = 0, cy0 = 0, cy1 = -1, cz1 = -1;
  auto begin_pass = [&]() {
    if (loose) {
      bx0 = bx1 = pass == 0 ? px[0] : pass == 1 ? px[1] : pass == 2 ? px[2] : px[3];
      by0 = by1 = pass == 0 ? py[0] : pass == 1 ? py[1] : pass == 2 ? py[2] : py[3];
      bz0 = bz1 = pass == 0 ? pz[0] : pass == 1 ? pz[1] : pass == 2 ? pz[2] : pz[3];
    }
    cy0 = max((int)floorf((by0 - search_f) * inv_edge), 0);
    cy1 = min((int)floorf((by1 + search_f) * inv_edge), g.ny - 1);
    z = max((int)floorf((bz0 - search_f) * inv_edge), 0);
    cz1 = min((int)floorf((bz1 + search_f) * inv_edge), g.nz - 1);
    y = cy0;
  };
  if (active) begin_pass();
  int cnt[4] = {0, 0, 0, 0};
  int ucnt = 0;

  while (__any_sync(0xffffffffu, active && (z <= cz1 || pass + 1 < npass))) {
    if (active && z > cz1 && pass + 1 < npass) {  // next member of a loose cluster
      pass++;
      live = 1u << pass;
      begin_pass();
    }
    uint32_t m = 0, m_end = 0;
    if (active && z <= cz1) {
      const float gy = fmaxf(fmaxf(y * edge - by1, by0 - (y + 1) * edge) - pad, 0.f);
      const float gz = fmaxf(fmaxf(z * edge - bz1, bz0 - (z + 1) * edge) - pad, 0.f);
      const float rem = hi_f - gy * gy - gz * gz;
      if (rem > 0.f) {
        const float w = sqrtf(rem) + pad;  // half chord of the search sphere along x
        const int xa = max((int)floorf((bx0 - w) * inv_edge), 0);
        const int xb = min((int)floorf((bx1 + w) * inv_edge), g.nx - 1);
        if (xa <= xb) {
          const int rowc = (z * g.ny + y) * g.nx;
          m = cell_start[rowc + xa] + lg;
          m_end = cell_start[rowc + xb + 1];
        }
      }
      if (++y > cy1) { y = cy0; z++; }
    }
    while (__any_sync(0xffffffffu, m < m_end)) {
      // lanes past the end of the range test a far-away dummy: no hit, no exact path
      float4 c32 = make_float4(3.0e18f, 0.f, 0.f, 0.f);
      if (m < m_end) c32 = sorted_pos32[m];
      const int j = __float_as_int(c32.w);
      // straight-line classification against the four members: two band tests per member
      //   sure hit  <=> margin < r2 < lo_f   (|r2 - mid_s| < hw_s)
      //   exact     <=> r2 < hi_f and not a sure hit (near the threshold, or near zero)
      bool hit[4];
      bool ex = false;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const float dx = px[r] - c32.x, dy = py[r] - c32.y, dz = pz[r] - c32.z;
        const float r2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        hit[r] = fabsf(r2f - mid_s) < hw_s;
        ex = ex || (r2f < hi_f && !hit[r]);
      }
      if (ex) {  // rare
        const double4 cj = sorted_pos[m];
#pragma unroll  // fully unrolled: no dynamic indexing of the register arrays
        for (int r = 0; r < 4; r++) {
          const float dx = px[r] - c32.x, dy = py[r] - c32.y, dz = pz[r] - c32.z;
          const float r2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          if (r2f < hi_f && !(fabsf(r2f - mid_s) < hw_s)) {
            double xi, yi, zi;
            load_pos<LAYOUT>(q, i0 + r, plane, xi, yi, zi);
            const double ddx = xi - cj.x, ddy = yi - cj.y, ddz = zi - cj.z;
            hit[r] = (j != (int)(i0 + r)) && (fma(ddz, ddz, fma(ddy, ddy, ddx * ddx)) < sl2);
          }
        }
      }
      if (half) {
#pragma unroll
        for (int r = 0; r < 4; r++) hit[r] = hit[r] && j > (int)(i0 + r);
      }
      if (live != 0xfu) {
#pragma unroll
        for (int r = 0; r < 4; r++) hit[r] = hit[r] && ((live >> r) & 1u);
      }
      const bool any = hit[0] || hit[1] || hit[2] || hit[3];
      if (FILL) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const unsigned b = __ballot_sync(0xffffffffu, hit[r]) & gbits;
          if (hit[r]) list[base[r] + cnt[r] + __popc(b & lt_mask)] = j;
          cnt[r] += __popc(b);
        }
        if (emit_cl) {
          const unsigned bu = __ballot_sync(0xffffffffu, any) & gbits;
          if (any) {
            const unsigned bits = (hit[0] ? 1u : 0u) | (hit[1] ? 2u : 0u) | (hit[2] ? 4u : 0u) | (hit[3] ? 8u : 0u);
            cl_list[cbase + ucnt + __popc(bu & lt_mask)] = (bits << 28) | (unsigned)j;
          }
          ucnt += __popc(bu);
        }
      } else {  // counting needs no ordering: per-lane tallies, reduced once at the end
#pragma unroll
        for (int r = 0; r < 4; r++) cnt[r] += hit[r] ? 1 : 0;
        ucnt += any ? 1 : 0;
      }
      m += GL;
    }
  }
  if (!FILL) {
#pragma unroll
    for (int s = GL / 2; s >= 1; s >>= 1) {
#pragma unroll
      for (int r = 0; r < 4; r++) cnt[r] += __shfl_xor_sync(0xffffffffu, cnt[r], s);
      ucnt += __shfl_xor_sync(0xffffffffu, ucnt, s);
    }
    if (nrows > 0) {
      if (lg < nrows) {
        const int mine = lg == 0 ? cnt[0] : lg == 1 ? cnt[1] : lg == 2 ? cnt[2] : cnt[3];
        nop[i0 + lg] = mine;
        if (mine > *(volatile int*)&tot->max_np) atomicMax(&tot->max_np, mine);
      }
      if (emit_cl && lg == 0) cl_cnt[c] = (uint32_t)ucnt;
    }
  }
}

__global__ void k_zero_u32(uint32_t* p, int64_t n, const int* n_dev) {
  if (n_dev) n = *n_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = 0u;
}

__global__ void k_zero_i32_rows(int32_t* p, int64_t pn) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pn;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = 0;
}

__global__ void k_finish_totals(lj_list_totals* tot, int64_t capacity, int pointer64) {
  if (tot->total > (unsigned long long)capacity) tot->overflow |= 1;
  if (!pointer64 && tot->total > 0xffffffffull) tot->overflow |= 2;
}

// ------------------------------------------------------------------ ELL / shuffle / check
template <bool PTR64>
__global__ void __launch_bounds__(256)
k_csr_to_ell(const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
             const void* __restrict__ pointer, int64_t pn, int max_np, int32_t* __restrict__ tl) {
  // thread per (row, k) with rows fastest: coalesced writes of the column-major table
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pn * (int64_t)max_np) return;
  const int64_t i = t % pn;
  const int k = (int)(t / pn);
  int v = 0;  // padding value 0, as thrust::fill(…, 0) in the reference (force_cuda.cu:230)
  if (k < nop[i]) v = list[row_offset<PTR64>(pointer, i) + k];
  tl[t] = v;
}

__global__ void k_max_np(const int32_t* __restrict__ nop, int64_t pn, int* out) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pn;
       i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, nop[i]);
  for (int s = 16; s >= 1; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// Fisher-Yates per row with a counter-based hash, one thread per row.
template <bool PTR64>
__global__ void __launch_bounds__(256)
k_shuffle_rows(int32_t* __restrict__ list, const int32_t* __restrict__ nop,
               const void* __restrict__ pointer, int64_t pn, uint32_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  int32_t* row = list + row_offset<PTR64>(pointer, i);
  const int n = nop[i];
  uint32_t state = mix32(seed ^ (uint32_t)i * 0x9e3779b9u);
  for (int k = n - 1; k > 0; k--) {
    state = mix32(state + 0x9e3779b9u);
    const int r = (int)(((uint64_t)state * (uint64_t)(k + 1)) >> 32);
    const int32_t t = row[k]; row[k] = row[r]; row[r] = t;
  }
}

template <bool PTR64>
__global__ void __launch_bounds__(256)
k_validate(const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
           const void* __restrict__ pointer, int64_t pn, int64_t npairs, int* bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < pn) {
    const int n = nop[t];
    const int64_t o = row_offset<PTR64>(pointer, t);
    if (n < 0 || n >= pn || o < 0 || o > npairs || o + n > npairs) atomicOr(bad, 1);
  }
  if (t < npairs) {
    const int j = list[t];
    if (j < 0 || j >= pn) atomicOr(bad, 2);
  }
}


// ------------------------------------------------------------------ cell-tile mirror ---
// LJ_LIST_TILES: besides the reference's CSR arrays the build emits the same list a second time,
// organised for the shared-memory force kernel (lj_force_celltile.cu): rows in cell order, each
// padded to a multiple of 8 entries, entries = 16-bit indices into the tile's staged region
// (geometry: lj_celltile.cuh).  Membership is decided by the same tests as k_search, so row s of
// the mirror is row order[s] of the CSR list as a set; the row lengths are taken from
// number_of_partners and cross-checked.
__global__ void k_tile_prepare(const grid_ext* __restrict__ ge, int64_t pn, int target_rows,
                               lj_tile_geom* __restrict__ tg) {
  const lj_grid_params g = ge->g;
  const double occ = fmax((double)pn / (double)g.ncell, 1.0e-3);  // mean particles per cell
  int tc0 = (int)((double)target_rows / occ + 0.5);
  tc0 = max(1, min(tc0, g.nx));
  const int ntx = (g.nx + tc0 - 1) / tc0;
  tg->tc = (g.nx + ntx - 1) / ntx;  // even split of the pencil
  tg->ntx = (g.nx + tg->tc - 1) / tg->tc;
  tg->nx = g.nx; tg->ny = g.ny; tg->nz = g.nz;
  tg->ntiles = tg->ntx * g.ny * g.nz;
  tg->max_rows = 0; tg->max_yrow = 0; tg->max_units = 0; tg->pad = 0;
  tg->ncols_active = 0; tg->pad2 = 0;
  tg->total_units = 0;
}

__global__ void __launch_bounds__(256)
k_tile_rows(int64_t pn, const float4* __restrict__ sorted_pos32, const int32_t* __restrict__ nop,
            int32_t* __restrict__ order, int32_t* __restrict__ cnt, uint32_t* __restrict__ units) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > pn) return;
  if (s == pn) { units[s] = 0u; return; }  // sentinel: the scan then yields off[pn] = total
  const int i = __float_as_int(sorted_pos32[s].w);
  const int c = nop[i];
  order[s] = i;
  cnt[s] = c;
  units[s] = (uint32_t)(c + 7) >> 3;
}

__global__ void __launch_bounds__(256)
k_tile_meta(int64_t pn, const int32_t* __restrict__ order, const int32_t* __restrict__ cnt,
            const uint32_t* __restrict__ off, int4* __restrict__ meta) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < pn) meta[s] = make_int4(cnt[s], (int)off[s], order[s], 0);
}

// One thread per (tx, Y = cy, cz).  Y-row table: the five pencil ranges {start, local base} of y-row
// (tx, Y, cz) + {0, length}.  Tile table: {first row, rows, first list unit, list units} and the
// region-local index of the tile's first row (it lives in the centre pencil: dy = 2, dz = 2).
__global__ void __launch_bounds__(128)
k_tile_table(const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ off,
             lj_tile_geom* __restrict__ tg, uint2* __restrict__ ytab, uint4* __restrict__ ttab,
             int32_t* __restrict__ col_flag, int phase) {  // phase 0: y-row table only, 1: tile table only, 2: both
  const lj_tile_geom g = *tg;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.ntiles) return;
  // tiles are numbered column by column: t = ((cz * ntx + tx) * ny + cy)
  const int cy = t % g.ny, tx = (t / g.ny) % g.ntx, cz = t / (g.ny * g.ntx);
  int xa, xb, rxa, rxb;
  tile_x_extent(tx * g.tc, g.tc, g.nx, xa, xb, rxa, rxb);
  uint32_t base = 0, st2 = 0, pb2 = 0;
  for (int dz = 0; dz < kTileYPencils; dz++) {
    uint32_t st, len;
    tile_pencil_range(cell_start, g.nx, g.ny, g.nz, cy, cz, rxa, rxb, dz, st, len);
    if (phase != 1) ytab[(size_t)t * kTileYTab + dz] = make_uint2(st, base);
    if (dz == 2) { st2 = st; pb2 = base; }
    base += len;
  }
  if (phase != 1) {
    ytab[(size_t)t * kTileYTab + 5] = make_uint2(0u, base);
    atomicMax(&tg->max_yrow, (int)base);
  }
  if (phase == 0) return;
  const int rowc = (cz * g.ny + cy) * g.nx;
  const uint32_t s0 = cell_start[rowc + xa], s1 = cell_start[rowc + xb + 1];
  ttab[(size_t)t * kTileTTab] = make_uint4(s0, s1 - s0, off[s0], off[s1] - off[s0]);
  ttab[(size_t)t * kTileTTab + 1] = make_uint4(pb2 + (s0 - st2), 0u, 0u, 0u);  // + 2 * cap_y at run time
  atomicMax(&tg->max_rows, (int)(s1 - s0));
  atomicMax(&tg->max_units, (int)(off[s1] - off[s0]));
  if (off[s1] != off[s0]) col_flag[t / g.ny] = 1;  // column (cz * ntx + tx) has work
}

// flags -> ascending list of the active columns (in place), count into the geometry.  One warp:
// a few thousand columns at most.
__global__ void k_tile_cols(int ncols, int32_t* __restrict__ cols, lj_tile_geom* __restrict__ tg) {
  const int lane = threadIdx.x;
  int count = 0;
  for (int base = 0; base < ncols; base += 32) {
    const int c = base + lane;
    const bool on = c < ncols && cols[c] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, on);  // every flag of this round is read before any write
    if (on) cols[count + __popc(b & ((1u << lane) - 1u))] = c;
    count += __popc(b);
    __syncwarp();
  }
  if (lane == 0) tg->ncols_active = count;
}

// FILL pass of the mirror: k_search in cell order, writing region-local 16-bit indices.  With
// PUBLIC it is the FILL pass of the reference-format list as well (each hit is stored twice: j
// into sorted_list, the local index into the mirror), which replaces the cluster FILL pass.
template <bool PUBLIC, bool PTR64>
__global__ void __launch_bounds__(256)
k_tile_fill(int64_t pn, const grid_ext* __restrict__ ge, const lj_tile_geom* __restrict__ tgp,
            const int32_t* __restrict__ cell_of, const uint32_t* __restrict__ cell_start,
            const double4* __restrict__ sorted_pos, const float4* __restrict__ sorted_pos32, double sl2,
            const int32_t* __restrict__ tl_cnt, const uint32_t* __restrict__ tl_off,
            const uint2* __restrict__ tab, uint16_t* __restrict__ tl_list, lj_list_totals* __restrict__ tot,
            const void* __restrict__ pointer, int32_t* __restrict__ list, int64_t capacity, int fake) {
  constexpr int GL = kSearchLanes;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GL;  // slot in cell order
  const int lane = threadIdx.x & 31;
  const int lg = lane % GL;
  const unsigned gbits = ((1u << GL) - 1u) << (lane - lg);
  const unsigned lt_mask = (1u << lane) - 1u;
  bool active = s < pn;
  float4 me32 = make_float4(0.f, 0.f, 0.f, 0.f);
  int want = 0;
  if (active) { me32 = sorted_pos32[s]; want = tl_cnt[s]; }
  const int i = __float_as_int(me32.w);
  active = active && want > 0;  // empty rows (also: rows outside the build's row range)
  const lj_grid_params g = ge->g;
  const int tc = tgp->tc, ntx = tgp->ntx;
  const float margin = ge->margin, sl2f = ge->sl2f, edge = ge->edge, pad = ge->pad;
  const float lo_f = sl2f - margin, hi_f = sl2f + margin;
  const int c = active ? cell_of[i] : 0;
  const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
  // y-row table of my column (tx, cz), row index Y in [0, ny)
  const uint2* __restrict__ mytab = tab + (size_t)((cz * ntx + cx / tc) * g.ny) * kTileYTab;
  const uint32_t cap_y = (uint32_t)((tgp->max_yrow + 8 + 1) & ~1);
  const size_t base = active ? (size_t)tl_off[s] * 8 : 0;
  int64_t pbase = 0;
  bool pub = PUBLIC && active;
  if (pub) {
    pbase = row_offset<PTR64>(pointer, i);
    if (pbase + want > capacity) {
      if (lg == 0) atomicOr(&tot->overflow, 1);
      pub = false;
    }
  }

  const float kInf = __int_as_float(0x7f800000);
  float gx2[5], gy2[5], gz2[5];
#pragma unroll
  for (int o = 0; o < 5; o++) {
    const int x = cx + o - 2, y = cy + o - 2, z = cz + o - 2;
    const float ax = fmaxf(fmaxf(x * edge - me32.x, me32.x - (x + 1) * edge) - pad, 0.f);
    const float ay = fmaxf(fmaxf(y * edge - me32.y, me32.y - (y + 1) * edge) - pad, 0.f);
    const float az = fmaxf(fmaxf(z * edge - me32.z, me32.z - (z + 1) * edge) - pad, 0.f);
    gx2[o] = (active && x >= 0 && x < g.nx) ? ax * ax : kInf;
    gy2[o] = (active && y >= 0 && y < g.ny) ? ay * ay : kInf;
    gz2[o] = (active && z >= 0 && z < g.nz) ? az * az : kInf;
  }
  int count = 0;
  double4 me = make_double4(0, 0, 0, 0);
  bool have_me = false;

#pragma unroll
  for (int dz = 0; dz < 5; dz++) {
#pragma unroll
    for (int dy = 0; dy < 5; dy++) {
      const float rem = hi_f - gz2[dz] - gy2[dy];
      const bool run = gx2[2] < rem;
      if (!__any_sync(0xffffffffu, run)) continue;
      uint32_t m = 0, m_end = 0, delta = 0;
      if (run) {
        const int xa = cx - ((gx2[0] < rem) ? 2 : (gx2[1] < rem) ? 1 : 0);
        const int xb = cx + ((gx2[4] < rem) ? 2 : (gx2[3] < rem) ? 1 : 0);
        const int rowc = ((cz + dz - 2) * g.ny + (cy + dy - 2)) * g.nx;
        m = cell_start[rowc + xa] + lg;
        m_end = cell_start[rowc + xb + 1];
        const uint2 e = mytab[(cy + dy - 2) * kTileYTab + dz];
        delta = (uint32_t)dy * cap_y + e.y - e.x;  // region-local index = sorted index + delta (mod 2^32)
      }
      while (__any_sync(0xffffffffu, m < m_end)) {
        bool hit = false;
        int j = 0;
        if (m < m_end) {
          const float4 c32 = sorted_pos32[m];
          j = __float_as_int(c32.w);
          const float dx = me32.x - c32.x, dy_ = me32.y - c32.y, dz_ = me32.z - c32.z;
          const float r2f = fmaf(dz_, dz_, fmaf(dy_, dy_, dx * dx));
          if (r2f < hi_f) {
            if (r2f < lo_f && r2f > margin) {
              hit = true;
            } else {
              if (!have_me) { me = sorted_pos[s]; have_me = true; }
              const double4 cj = sorted_pos[m];
              const double ddx = me.x - cj.x, ddy = me.y - cj.y, ddz = me.z - cj.z;
              hit = (j != i) && (fma(ddz, ddz, fma(ddy, ddy, ddx * ddx)) < sl2);
            }
          }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit) & gbits;
        const int pos = count + __popc(ballot & lt_mask);
        if (hit && pos < want) {
          // fake (LJ_TILE_FAKE, timing experiment only): a bank-conflict-free index pattern
          tl_list[base + pos] = fake ? (uint16_t)((pos & 7) + ((s & 1) << 3) + (((pos >> 3) & 3) << 4))
                                     : (uint16_t)(m + delta);
          if (pub) list[pbase + pos] = j;
        }
        count += __popc(ballot);
        m += GL;
      }
    }
  }
  if (active) {
    if (count != want) {  // cannot happen unless the CSR arrays were not built from these positions
      if (lg == 0) atomicOr(&tot->overflow, 8);
    }
    const uint16_t dummy = (uint16_t)(cap_y - 1);  // last record of the dy = 0 ring slot: far-away point
    const int padded = ((want + 7) >> 3) << 3;
    for (int k = min(count, want) + lg; k < padded; k += GL) tl_list[base + k] = dummy;
  }
}

// Column selection of a PART launch (lj_force_step_part): the active columns (tx, cz) whose five
// stencil cell layers cz-2 .. cz+2 held no particle outside the list's row range at build time
// (INTERIOR: their rows depend on positions inside the row range only -- in a z-slab run, on no
// ghost), or the others (BOUNDARY).  Run by block 0 of the permute kernel that precedes the force
// kernel; the order of the compacted list does not matter (units are dealt dynamically).
struct tile_part { int part; int64_t r0, r1; const int32_t* cols; int ncols_all; const int32_t* zflag; int32_t* cols_sel; };

__device__ __forceinline__ void tile_select_columns(const tile_part& tp, const lj_tile_geom* __restrict__ tg,
                                                     lj_tile_geom* __restrict__ tg_out) {
  __shared__ int n_sel;
  if (threadIdx.x == 0) n_sel = 0;
  __syncthreads();
  const int ntx = tg->ntx, nz = tg->nz;
  const bool all_cols = tg->ncols_active <= 0 || tg->ncols_active >= ntx * nz;
  const int ncols = all_cols ? ntx * nz : tg->ncols_active;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    const int col = all_cols ? c : tp.cols[c];
    const int cz = col / ntx;
    bool touched = false;
    for (int z = max(cz - 2, 0); z <= min(cz + 2, nz - 1); z++) touched |= tp.zflag[z] != 0;
    if ((tp.part == 1) == !touched) tp.cols_sel[atomicAdd(&n_sel, 1)] = col;
  }
  __syncthreads();
  if (threadIdx.x == 0) tg_out->pad2 = n_sel;  // read by the force kernel that follows
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_tile_permute(const void* __restrict__ q, int64_t plane, const int32_t* __restrict__ order,
               int64_t pn, double2* __restrict__ qxy, double* __restrict__ qz, lj_tile_geom* __restrict__ tg,
               const tile_part tp) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) tg->pad = 0;  // the force kernel that follows deals its work units from here
  if (tp.part != 0 && blockIdx.x == 0) tile_select_columns(tp, tg, tg);
  if (s >= pn) return;
  const int o = order[s];
  if (tp.part != 0 && ((o >= tp.r0 && o < tp.r1) != (tp.part == 1))) return;  // INTERIOR: rows of the range; BOUNDARY: the others
  double x, y, z;
  load_pos<LAYOUT>(q, o, plane, x, y, z);
  qxy[s] = make_double2(x, y);  // two planes: the force kernel reads {x,y} with LDS.128, z with LDS.64
  qz[s] = z;
}

// the same for the mixed-precision kernel: {x, y, z} in counts modulo 2^32 (lj_fx_frame), .w = the
// original index (the kernel re-decides borderline cutoff cases from the caller's FP64 positions)
template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_tile_permute_fx(const void* __restrict__ q, int64_t plane, const int32_t* __restrict__ order,
                  int64_t pn, double scale, int4* __restrict__ qfx, lj_tile_geom* __restrict__ tg,
                  const tile_part tp) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s == 0) tg->pad = 0;
  if (tp.part != 0 && blockIdx.x == 0) tile_select_columns(tp, tg, tg);
  if (s >= pn) return;
  const int o = order[s];
  if (tp.part != 0 && ((o >= tp.r0 && o < tp.r1) != (tp.part == 1))) return;
  double x, y, z;
  load_pos<LAYOUT>(q, o, plane, x, y, z);
  qfx[s] = make_int4((int)(uint32_t)__double2ll_rn(x * scale), (int)(uint32_t)__double2ll_rn(y * scale),
                     (int)(uint32_t)__double2ll_rn(z * scale), o);
}

// cell layers (z index) that hold a particle outside the list's row range [r0, r1)
__global__ void __launch_bounds__(256)
k_tile_zflag(int64_t pn, int64_t r0, int64_t r1, const int32_t* __restrict__ cell_of,
             const lj_tile_geom* __restrict__ tg, int32_t* __restrict__ zflag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn || (i >= r0 && i < r1)) return;
  zflag[cell_of[i] / (tg->nx * tg->ny)] = 1;
}

// ====================================================================== tile engine ======
// ONE search per build (round 2).  Round 1 searched twice: k_search_cluster counted (0.75 ms at
// N = 1M) and k_tile_fill searched again to write the list and its mirror (1.68 ms).  Here the
// COUNT pass runs per tile with the tile's 25-pencil region staged in shared memory, and records
// for every row and pencil WHICH candidates hit as a 64-bit mask (bit b = record b of the row's
// window = the five x-cells around its own cell).  After the scans, k_tile_replay walks the masks
// and writes both lists without looking at a single position.  Membership is decided by the same
// tests as k_search (FP32 on origin-shifted coordinates, the exact FP64 fma chain inside the error
// band); a row's entries are emitted pencil by pencil in the order kTePerm, cell order within a pencil.
constexpr int kTeThreads = 192;  // replay / translate: six warps, four rows of eight lanes each
constexpr int kTeLanes = 8;      // lanes per row (pair), four per warp
constexpr int kTeRowKx = 256;    // replay: rows of a tile whose region x-cell is tabulated (tiles hold ~56-72)
constexpr int kTeRowCap = 256;   // replay: entries of a row staged in shared memory at most (longer rows are written directly);
                                 // the launch sizes the staging area for the longest row of THIS list (occupancy: the pass is
                                 // latency-bound, 256 entries per row = 5 CTAs per SM, 176 = 6: 1.40 -> 1.29 ms per build)

// replay: bytes a row group takes in the staging area (rowcap 4-byte + rowcap 2-byte entries, padded: see the kernel)
__host__ __device__ inline uint32_t te_group_stride(int rowcap) {
  uint32_t w = ((uint32_t)rowcap * 6u + 3u) / 4u;
  w += (8u - (w & 31u)) & 31u;
  return w * 4u;
}

struct te_pencil { uint32_t st, base, len; int rowc; };  // cell-order start, region-local index of its first record,
                                                          // records, first cell of the (y,z) row (-1: outside the grid)

// pencil p = dz * 5 + dy of tile t = (tx, cy, cz), from the y-row table (the y-rows cy-2 .. cy+2 of a column
// are the table rows t-2 .. t+2); window tables: xoff[p][k] = first record (relative to the pencil's first
// record) of region x-cell k, k = 0 .. ncx
__device__ __forceinline__ void te_setup(const lj_tile_geom& g, int t, uint32_t cap_y, const uint32_t* __restrict__ cell_start,
                                         const uint2* __restrict__ ytab, te_pencil* __restrict__ pen, int* __restrict__ xoff,
                                         int ncx1, int& ncx, uint32_t& s0, uint32_t& ns) {
  const int cy = t % g.ny, tx = (t / g.ny) % g.ntx, cz = t / (g.ny * g.ntx);
  int xa, xb, rxa, rxb;
  tile_x_extent(tx * g.tc, g.tc, g.nx, xa, xb, rxa, rxb);
  ncx = rxb - rxa + 1;
  // one global round trip: the y-row table entries (threads 0 .. 24) and the cell starts of the 25 window
  // tables are loaded side by side, the pencil's first record is subtracted once both have arrived
  if (threadIdx.x < 25) {
    const int p = threadIdx.x, dz = p / 5, dy = p % 5;
    const int Y = cy + dy - 2, z = cz + dz - 2;
    te_pencil e;
    e.st = 0; e.base = 0; e.len = 0;
    e.rowc = (Y < 0 || Y >= g.ny || z < 0 || z >= g.nz) ? -1 : (z * g.ny + Y) * g.nx;
    if (Y >= 0 && Y < g.ny) {
      const uint2* row = ytab + (size_t)(t + dy - 2) * kTileYTab;
      const uint2 a = row[dz], nx = row[dz + 1];   // entry 5 = {0, length of the y-row}
      e.st = a.x; e.base = (uint32_t)dy * cap_y + a.y; e.len = nx.y - a.y;
    }
    pen[p] = e;
  }
  for (int idx = threadIdx.x; idx < 25 * ncx1; idx += blockDim.x) {
    const int p = idx / ncx1, k = idx - p * ncx1;
    const int Y = cy + p % 5 - 2, z = cz + p / 5 - 2;
    const bool inside = Y >= 0 && Y < g.ny && z >= 0 && z < g.nz && k <= ncx;
    xoff[idx] = inside ? (int)cell_start[(z * g.ny + Y) * g.nx + rxa + k] : -1;
  }
  const int rowc = (cz * g.ny + cy) * g.nx;
  s0 = cell_start[rowc + xa];
  ns = cell_start[rowc + xb + 1] - s0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 25 * ncx1; idx += blockDim.x) {
    const int v = xoff[idx];
    xoff[idx] = v >= 0 ? v - (int)pen[idx / ncx1].st : 0;
  }
  __syncthreads();
}

// region x-cell of the tile row with offset `rel` into the centre pencil
__device__ __forceinline__ int te_row_cell(const int* __restrict__ xoff_c, int ncx, uint32_t rel) {
  int kx = 0;
  while (kx + 1 < ncx && (uint32_t)xoff_c[kx + 1] <= rel) kx++;
  return kx;
}

// COUNT pass, second form: ONE THREAD PER (row, pencil).  A tile of ~56 rows has 1400 such jobs, lanes are
// consecutive rows of one pencil, so the lanes of a warp read the same few candidate records (rows that are
// neighbours in cell order share their window: a multicast shared-memory load), nobody waits for a slower
// lane (the trip count is the warp's largest window, rounded to 8) and there is no per-pencil set-up shared
// by only two rows: 11 instructions per test instead of 16 and no 90-instruction pencil prologue per eight
// rows -- the count pass executes 2.5x fewer instructions than k_tile_count.  Same tests, same windows
// (the window of a row is that of its PAIR (2k, 2k + 1): k_tile_replay recomputes it), mask bit b = window
// record b, masks stored pencil-major (masks[p * pn + s]: a warp writes 256 contiguous bytes).
constexpr int kTcThreads = 256;
constexpr int kTcRowChunk = 128;  // rows of a tile worked on at a time (tiles hold ~56, ~72 wide)

#define LJ_TC_TEST(H, U, J)                                                                                         \
  {                                                                                                                 \
    const float4 c = cand[(J)];                                                                                     \
    const float dx = me.x - c.x, dy = me.y - c.y, dz = me.z - c.z;                                                  \
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));                                                           \
    asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(H) : "f"(r2), "f"(lo_f), "n"(1u << ((J) & 31))); \
    asm("{\n.reg .pred p;\nsetp.lt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(U) : "f"(r2), "f"(hi_f), "n"(1u << ((J) & 31))); \
  }
#define LJ_TC_GROUP(H, U, K)                                                                                        \
  if (nmax > (K)) {                                                                                                 \
    LJ_TC_TEST(H, U, (K) + 0) LJ_TC_TEST(H, U, (K) + 1) LJ_TC_TEST(H, U, (K) + 2) LJ_TC_TEST(H, U, (K) + 3)          \
    LJ_TC_TEST(H, U, (K) + 4) LJ_TC_TEST(H, U, (K) + 5) LJ_TC_TEST(H, U, (K) + 6) LJ_TC_TEST(H, U, (K) + 7)          \
  }

__global__ void __launch_bounds__(kTcThreads, 4)
k_tile_count2(int64_t pn, int64_t r0, int64_t r1, const grid_ext* __restrict__ ge, const lj_tile_geom* __restrict__ tgp,
              const uint32_t* __restrict__ cell_start, const uint2* __restrict__ ytab,
              const double4* __restrict__ sorted_pos, const float4* __restrict__ sorted_pos32, double sl2,
              int ncx1, int32_t* __restrict__ nop, int32_t* __restrict__ tl_order, int32_t* __restrict__ tl_cnt,
              uint32_t* __restrict__ tl_units, unsigned long long* __restrict__ masks, lj_list_totals* __restrict__ tot) {
  extern __shared__ __align__(16) unsigned char te_smem[];
  const lj_tile_geom g = *tgp;
  const uint32_t cap_y = (uint32_t)((g.max_yrow + 8 + 1) & ~1);
  float4* reg = reinterpret_cast<float4*>(te_smem);  // 5 * cap_y records + 64 of slack (windows are read in
                                                      // groups of 8: a read past the region's end stays inside)
  te_pencil* pen = reinterpret_cast<te_pencil*>(reg + 5 * cap_y + 64);
  int* xoff = reinterpret_cast<int*>(pen + 25);
  int* rowcnt = xoff + 25 * ncx1;
  unsigned char* rowkx = reinterpret_cast<unsigned char*>(rowcnt + kTcRowChunk);
  __shared__ int blk_max;
  if (threadIdx.x == 0) blk_max = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) tl_units[pn] = 0u;  // sentinel: the scan then yields off[pn] = total
  int ncx;
  uint32_t s0, ns;
  te_setup(g, blockIdx.x, cap_y, cell_start, ytab, pen, xoff, ncx1, ncx, s0, ns);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ns == 0) return;  // (uniform) an empty tile
  for (int p = warp; p < 25; p += kTcThreads / 32) {  // stage the region: a warp per pencil
    const te_pencil e = pen[p];
    float4* dst = reg + e.base;
    for (uint32_t k = lane; k < e.len; k += 32) dst[k] = sorted_pos32[e.st + k];
  }
  const float margin = ge->margin, sl2f = ge->sl2f;
  const float lo_f = sl2f - margin, hi_f = sl2f + margin;
  const te_pencil pc = pen[12];  // the centre pencil holds the tile's rows
  const int* xoff_c = xoff + 12 * ncx1;
  int my_max = 0;
  for (uint32_t rbase = 0; rbase < ns; rbase += kTcRowChunk) {
    const int nr = (int)min((uint32_t)kTcRowChunk, ns - rbase);
    if (rbase > 0) __syncthreads();  // the previous chunk's counts have been read
    for (int r = threadIdx.x; r < nr; r += kTcThreads) {
      rowkx[r] = (unsigned char)te_row_cell(xoff_c, ncx, s0 + rbase + r - pc.st);
      rowcnt[r] = 0;
    }
    __syncthreads();  // (first chunk: the region is staged as well)
    const int njobs = nr * 25;
    for (int job0 = warp * 32; job0 < njobs; job0 += kTcThreads) {
      const int job = job0 + lane;
      const bool live = job < njobs;
      const int p = live ? job / nr : 0, r = live ? job - p * nr : 0;
      const uint32_t rr = rbase + (uint32_t)r;
      const uint32_t rp = (rr ^ 1u) < ns ? (rr ^ 1u) : rr;  // the other row of the pair (same chunk: the chunk is even)
      const uint32_t s = s0 + rr, rel = s - pc.st;
      const int kx = rowkx[r], kxp = rowkx[rp - rbase];
      const int k_lo = max(min(kx, kxp) - 2, 0), k_hi = min(max(kx, kxp) + 2, ncx - 1);
      const te_pencil e = pen[p];
      const float4 me = reg[pc.base + rel];
      const int ia = __float_as_int(me.w);
      const bool in = live && ia >= r0 && ia < r1;
      const int w0 = xoff[p * ncx1 + k_lo];
      int nw = (in && e.rowc >= 0) ? xoff[p * ncx1 + k_hi + 1] - w0 : 0;
      if (nw > 64) { atomicOr(&tot->overflow, 16); nw = 64; }
      const int nmax = __reduce_max_sync(0xffffffffu, nw);
      const float4* __restrict__ cand = reg + (e.base + (uint32_t)w0);
      // h: r2 < lo (a hit for sure), u: r2 < hi (a hit or inside the FP32 error band), one predicated OR each
      unsigned h0 = 0, u0 = 0, h1 = 0, u1 = 0;
      LJ_TC_GROUP(h0, u0, 0) LJ_TC_GROUP(h0, u0, 8) LJ_TC_GROUP(h0, u0, 16) LJ_TC_GROUP(h0, u0, 24)
      if (nmax > 32) {
        LJ_TC_GROUP(h1, u1, 32) LJ_TC_GROUP(h1, u1, 40) LJ_TC_GROUP(h1, u1, 48) LJ_TC_GROUP(h1, u1, 56)
      }
      const unsigned long long lm = nw >= 64 ? ~0ull : (1ull << nw) - 1ull;  // records past the window do not count
      unsigned long long ha = ((unsigned long long)h1 << 32 | h0) & lm;
      unsigned long long band = ((unsigned long long)u1 << 32 | u0) & lm & ~ha;
      if (p == 12 && nw > 0) {  // a row is not its own neighbour
        const unsigned long long self = 1ull << ((int)rel - w0);
        ha &= ~self; band &= ~self;
      }
      while (band) {  // rare (about one candidate in 5000): the exact FP64 test
        const int k = __ffsll((long long)band) - 1;
        band &= band - 1;
        const uint32_t m = e.st + (uint32_t)(w0 + k);
        const double4 cj = sorted_pos[m], md = sorted_pos[s];
        const double dx = md.x - cj.x, dy = md.y - cj.y, dz = md.z - cj.z;
        if (fma(dz, dz, fma(dy, dy, dx * dx)) < sl2) ha |= 1ull << k;
      }
      if (in) {
        masks[(size_t)p * (size_t)pn + s] = ha;
        const int c = __popcll(ha);
        if (c) atomicAdd(&rowcnt[r], c);
      }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < nr; r += kTcThreads) {
      const uint32_t s = s0 + rbase + (uint32_t)r;
      const int ia = __float_as_int(reg[pc.base + (s - pc.st)].w), c = rowcnt[r];
      tl_order[s] = ia; tl_cnt[s] = c; tl_units[s] = (uint32_t)(c + 7) >> 3; nop[ia] = c;
      my_max = max(my_max, c);
    }
  }
  if (my_max > 0) atomicMax(&blk_max, my_max);
  __syncthreads();
  if (threadIdx.x == 0 && blk_max > *(volatile int*)&tot->max_np) atomicMax(&tot->max_np, blk_max);
}
#undef LJ_TC_GROUP
#undef LJ_TC_TEST

// REPLAY pass: masks -> mirror entries (16-bit region-local indices) and the public list (original
// indices), both at the offsets the scans produced.  No positions are read.  A row's entries are
// expanded into shared memory (each lane its pencils, scattered) and written out by the row's eight
// lanes in runs of eight consecutive entries: 16 contiguous bytes of the mirror, 32 of the list.
// replay: the order in which a row's 25 pencils (p = dz * 5 + dy) are expanded, by expected number of hits:
// (0,0); the four with |dy| + |dz| = 1; (+-1,+-1); (+-2,0), (0,+-2); the eight (+-2,+-1), (+-1,+-2); the corners
__constant__ unsigned char kTePerm[32] = {12, 7,  11, 13, 17, 6,  8,  16,  18, 2,  10, 14, 22, 1,  3,  5,
                                          9,  15, 19, 21, 23, 0,  4,  20,  24, 255, 255, 255, 255, 255, 255, 255};

#ifndef LJ_TE_MINBLOCKS
#define LJ_TE_MINBLOCKS 7  // 48 registers: seven CTAs per SM (the pass waits on dependent global loads)
#endif
template <bool PTR64>
__global__ void __launch_bounds__(kTeThreads, LJ_TE_MINBLOCKS)
k_tile_replay(int64_t pn, const lj_tile_geom* __restrict__ tgp, const uint32_t* __restrict__ cell_start,
              const uint2* __restrict__ ytab, int ncx1, const int32_t* __restrict__ tl_order,
              const int32_t* __restrict__ tl_cnt, const uint32_t* __restrict__ tl_off,
              const unsigned long long* __restrict__ masks, uint16_t* __restrict__ tl_list,
              const void* __restrict__ pointer, int32_t* __restrict__ list, int64_t capacity, int rowcap,
              lj_list_totals* __restrict__ tot) {
  extern __shared__ __align__(16) unsigned char te_smem[];
  const lj_tile_geom g = *tgp;
  const uint32_t cap_y = (uint32_t)((g.max_yrow + 8 + 1) & ~1);
  te_pencil* pen = reinterpret_cast<te_pencil*>(te_smem);
  int* xoff = reinterpret_cast<int*>(pen + 25);
  // per row group: rowcap x {cell-order index of j (4 B)} then rowcap x {region-local index (2 B)}
  unsigned char* stage = reinterpret_cast<unsigned char*>(xoff + 25 * ncx1);
  int ncx;
  uint32_t s0, ns;
  te_setup(g, blockIdx.x, cap_y, cell_start, ytab, pen, xoff, ncx1, ncx, s0, ns);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lg = lane % kTeLanes, gi = lane / kTeLanes;
  const unsigned gmask = 0xffu << (gi * kTeLanes);
  const te_pencil pc = pen[12];
  const int* xoff_c = xoff + 12 * ncx1;
  const uint16_t dummy = (uint16_t)(cap_y - 1);  // last record of the dy = 0 ring slot: far-away point
  // region x-cell of every row, once per tile (one thread per row) instead of twice per row group by all of
  // its lanes: the linear search was 18 % of the kernel's instructions
  __shared__ unsigned char rowkx[kTeRowKx];
  const uint32_t ntab = ncx <= 256 ? min(ns, (uint32_t)kTeRowKx) : 0u;  // (an x-cell fits in a byte)
  for (uint32_t r = threadIdx.x; r < ntab; r += kTeThreads)
    rowkx[r] = (unsigned char)te_row_cell(xoff_c, ncx, s0 + r - pc.st);
  __syncthreads();
  // stride of a row group: 8 words mod 32, so that the four groups of a warp (same k, eight consecutive words
  // each) read four different sets of banks in the write-out loop
  const uint32_t gstride = te_group_stride(rowcap);
  uint32_t* sm_m = reinterpret_cast<uint32_t*>(stage + (size_t)(warp * 4 + gi) * gstride);
  uint16_t* sm_l = reinterpret_cast<uint16_t*>(sm_m + rowcap);
  for (uint32_t rb = warp * 4; rb < ns; rb += (kTeThreads / 32) * 4) {
    const uint32_t r = rb + gi;
    const bool valid = r < ns;
    const uint32_t s = s0 + (valid ? r : 0u);
    // the row's masks are requested together with its count: one round trip to global memory instead of two
    unsigned long long mk[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      const int p = kTePerm[q4 * kTeLanes + lg];
      mk[q4] = (valid && p < 25) ? masks[(size_t)p * (size_t)pn + s] : 0ull;  // (pencil-major: four rows of a warp = one sector)
    }
    const int want = valid ? tl_cnt[s] : 0;
    if (want == 0) continue;  // (whole groups: the eight lanes of a row agree)
    // the window of the COUNT pass: shared with the other row of the pair (2k, 2k + 1)
    const uint32_t rp = (r ^ 1u) < ns ? (r ^ 1u) : r;
    const uint32_t rv = valid ? r : 0u;
    const int kx = rv < ntab ? rowkx[rv] : te_row_cell(xoff_c, ncx, s - pc.st);
    const int kxp = rp < ntab ? rowkx[rp] : te_row_cell(xoff_c, ncx, s0 + rp - pc.st);
    const int* xlo = xoff + max(min(kx, kxp) - 2, 0);
    const size_t base = (size_t)tl_off[s] * 8;
    const int i = tl_order[s];
    const int64_t pbase = row_offset<PTR64>(pointer, i);
    bool pub = true;
    if (pbase + want > capacity) {
      if (lg == 0) atomicOr(&tot->overflow, 1);
      pub = false;
    }
    const bool staged = want <= rowcap;
    // lane lg owns the pencils kTePerm[lg], [lg + 8], [lg + 16] (and [24] for lg = 0); entries are emitted in
    // that order.  The eight pencils of a step hold about the same number of hits (the step costs its longest
    // mask): the centre pencil and its nearest eight first, the corners last.
    int before = 0;  // entries of all earlier pencils
    int off[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      const int c = __popcll(mk[q4]);
      int inc = c;  // inclusive scan over the eight lanes of the row
#pragma unroll
      for (int d = 1; d < kTeLanes; d <<= 1) {
        const int v = __shfl_up_sync(gmask, inc, d, kTeLanes);
        if (lg >= d) inc += v;
      }
      off[q4] = before + inc - c;
      before += __shfl_sync(gmask, inc, kTeLanes - 1, kTeLanes);
    }
    if (before != want) { if (lg == 0) atomicOr(&tot->overflow, 8); continue; }  // cannot happen
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
      const int p = kTePerm[q4 * kTeLanes + lg];
      if (p >= 25) continue;
      const te_pencil e = pen[p];
      const int w0 = xlo[p * ncx1];
      const uint32_t lbase = e.base + (uint32_t)w0, mbase = e.st + (uint32_t)w0;
      int o = off[q4];
      // the two 32-bit halves separately: windows rarely hold more than 32 records, and 32-bit
      // find-first-set / clear-lowest are one instruction each where the 64-bit forms are four
#pragma unroll
      for (int half = 0; half < 2; half++) {
        unsigned m = half ? (unsigned)(mk[q4] >> 32) : (unsigned)mk[q4];
        const uint32_t lb = lbase + 32u * half, mb = mbase + 32u * half;
        while (m) {
          const int b = __ffs((int)m) - 1;
          m &= m - 1;
          if (staged) { sm_m[o] = mb + b; sm_l[o] = (uint16_t)(lb + b); }
          else {
            tl_list[base + o] = (uint16_t)(lb + b);
            if (pub) list[pbase + o] = tl_order[mb + b];
          }
          o++;
        }
      }
    }
    const int padded = ((want + 7) >> 3) << 3;
    if (staged) {
      __syncwarp(gmask);
#pragma unroll 4
      for (int k = lg; k < padded; k += kTeLanes) {  // (unrolled: four gathers of tl_order in flight per lane)
        const bool real = k < want;
        tl_list[base + k] = real ? sm_l[k] : dummy;
        if (pub && real) list[pbase + k] = tl_order[sm_m[k]];
      }
      __syncwarp(gmask);
    } else {
      for (int k = want + lg; k < padded; k += kTeLanes) tl_list[base + k] = dummy;
    }
  }
}

// ---- lj_list_mirror: the mirror of a list the CALLER built ---------------------------------
__global__ void __launch_bounds__(256)
k_slot_of(int64_t pn, const int32_t* __restrict__ order, int32_t* __restrict__ slot_of) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < pn) slot_of[order[s]] = (int32_t)s;
}

// One CTA per tile, eight lanes per row: entry k of row i (the caller's order) -> region-local index
// of j in i's tile.  A row with an entry outside the 25 pencils / the region's x-range is flagged
// (row_flag[i] = 1, its mirror row gets length 0) and left to the per-row kernel.
template <bool PTR64>
__global__ void __launch_bounds__(kTeThreads)
k_tile_translate(int64_t pn, const lj_tile_geom* __restrict__ tgp, const uint32_t* __restrict__ cell_start,
                 const uint2* __restrict__ ytab, int ncx1, const int32_t* __restrict__ cell_of,
                 const int32_t* __restrict__ slot_of, const int32_t* __restrict__ tl_order,
                 const int32_t* __restrict__ tl_cnt, const uint32_t* __restrict__ tl_off, int4* __restrict__ meta,
                 uint16_t* __restrict__ tl_list, const void* __restrict__ pointer, const int32_t* __restrict__ list,
                 unsigned char* __restrict__ row_flag, lj_list_totals* __restrict__ tot) {
  extern __shared__ __align__(16) unsigned char te_smem[];
  const lj_tile_geom g = *tgp;
  const uint32_t cap_y = (uint32_t)((g.max_yrow + 8 + 1) & ~1);
  te_pencil* pen = reinterpret_cast<te_pencil*>(te_smem);
  int* xoff = reinterpret_cast<int*>(pen + 25);
  int ncx;
  uint32_t s0, ns;
  const int t = blockIdx.x;
  te_setup(g, t, cap_y, cell_start, ytab, pen, xoff, ncx1, ncx, s0, ns);
  const int cy = t % g.ny, tx = (t / g.ny) % g.ntx, cz = t / (g.ny * g.ntx);
  int xa, xb, rxa, rxb;
  tile_x_extent(tx * g.tc, g.tc, g.nx, xa, xb, rxa, rxb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lg = lane % kTeLanes, gi = lane / kTeLanes;
  const unsigned gmask = 0xffu << (gi * kTeLanes);
  const uint16_t dummy = (uint16_t)(cap_y - 1);
  for (uint32_t rb = warp * 4; rb < ns; rb += (kTeThreads / 32) * 4) {
    const uint32_t r = rb + gi;
    if (r >= ns) continue;  // (whole groups)
    const uint32_t s = s0 + r;
    const int want = tl_cnt[s];
    const int i = tl_order[s];
    const size_t base = (size_t)tl_off[s] * 8;
    const int64_t pbase = row_offset<PTR64>(pointer, i);
    bool outside = false;
    for (int k = lg; k < want; k += kTeLanes) {
      const int j = list[pbase + k];
      uint16_t L = dummy;
      bool ok = j >= 0 && j < pn;
      if (ok) {
        const int c = cell_of[j];
        const int jx = c % g.nx, jy = (c / g.nx) % g.ny, jz = c / (g.nx * g.ny);
        const int dy = jy - cy + 2, dz = jz - cz + 2;
        ok = dy >= 0 && dy < 5 && dz >= 0 && dz < 5 && jx >= rxa && jx <= rxb;
        if (ok) {
          const te_pencil e = pen[dz * 5 + dy];
          L = (uint16_t)(e.base + ((uint32_t)slot_of[j] - e.st));
        }
      }
      outside |= !ok;
      tl_list[base + k] = L;
    }
    const int padded = ((want + 7) >> 3) << 3;
    for (int k = want + lg; k < padded; k += kTeLanes) tl_list[base + k] = dummy;
    outside = __any_sync(gmask, outside);
    if (lg == 0) {
      row_flag[i] = outside ? 1 : 0;
      if (outside) {
        meta[s].x = 0;  // the cell-tile kernel skips the row
        atomicAdd(&tot->cl_total, 1ull);  // (reused as the count of rows left to the per-row kernel)
      }
    }
  }
}

int64_t blocks_for(int64_t n, int tb) { return (n + tb - 1) / tb; }

}  // namespace

// --------------------------------------------------------------------------- scratch ---
int lj_scratch_reserve(lj_ctx* ctx, int64_t pn, cudaStream_t st) {
  if (!ctx->totals) {
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->bbox, 8 * sizeof(double), ctx->pool, st));
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->grid, 256, ctx->pool, st));
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->totals, sizeof(lj_list_totals) + 16, ctx->pool, st));
    LJ_CUDA(ctx, cudaHostAlloc((void**)&ctx->totals_host, sizeof(lj_list_totals) + 16, cudaHostAllocDefault));
  }
  if (pn <= ctx->scratch_pn) return LJ_OK;
  void* olds[] = {ctx->cell_of, ctx->cell_slot, ctx->cell_count, ctx->cell_start,
                  ctx->sorted_pos, ctx->sorted_tmp, ctx->scan_tmp, ctx->q32};
  for (void* o : olds)
    if (o) LJ_CUDA(ctx, cudaFreeAsync(o, st));
  ctx->q32 = nullptr; ctx->q32_len = 0;
  const int64_t cells = pn < 32768 ? 32768 : pn;
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cell_of, sizeof(int32_t) * pn, ctx->pool, st));
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cell_slot, sizeof(int32_t) * pn, ctx->pool, st));
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cell_count, sizeof(uint32_t) * (cells + 1), ctx->pool, st));
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cell_start, sizeof(uint32_t) * (cells + 1), ctx->pool, st));
  // sorted_pos (double4) followed by sorted_pos32 (float4)
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->sorted_pos, (sizeof(double4) + sizeof(float4)) * pn, ctx->pool, st));
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->sorted_tmp, sizeof(int32_t) * pn, ctx->pool, st));
  ctx->scan_tmp_len = blocks_for(cells > pn ? cells : pn, kScanTile) + 1;
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->scan_tmp, sizeof(unsigned long long) * ctx->scan_tmp_len, ctx->pool, st));
  ctx->scratch_pn = pn;
  ctx->scratch_cells = cells;
  return LJ_OK;
}

// bounding box of q into ctx->bbox (6 ordered-encoded doubles); shared with the mixed kernels
int lj_bbox_launch(lj_ctx* ctx, const void* q, int layout, int64_t pn, int64_t plane,
                   lj_list_totals* reset_totals, cudaStream_t st) {
  unsigned long long* bb = reinterpret_cast<unsigned long long*>(ctx->bbox);
  const int nb = (int)(blocks_for(pn, 1024) < 8 * ctx->sm_count ? (blocks_for(pn, 1024) > 0 ? blocks_for(pn, 1024) : 1) : 8 * ctx->sm_count);
  k_bbox_init<<<1, 32, 0, st>>>(bb, reset_totals);
  LJ_LAUNCHED(ctx);
  switch (layout) {
    case LJ_AOS_D3: k_bbox<LJ_AOS_D3><<<nb, 256, 0, st>>>(q, pn, plane, bb); break;
    case LJ_AOS_D4: k_bbox<LJ_AOS_D4><<<nb, 256, 0, st>>>(q, pn, plane, bb); break;
    case LJ_AOS_F4: k_bbox<LJ_AOS_F4><<<nb, 256, 0, st>>>(q, pn, plane, bb); break;
    case LJ_AOS_F3: k_bbox<LJ_AOS_F3><<<nb, 256, 0, st>>>(q, pn, plane, bb); break;
    default: k_bbox<LJ_SOA_D><<<nb, 256, 0, st>>>(q, pn, plane, bb); break;
  }
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

// --------------------------------------------------------------------------- build -----
template <int LAYOUT>
static int cluster_fill(lj_ctx* ctx, const lj_list_args* a, cudaStream_t st, bool emit);

// *deferred (in: the caller would like to run the FILL pass itself; out: it has to): the COUNT pass
// and the scans are done, sorted_list is still unwritten.
// bounding box -> cell grid (on the device) -> counting sort of the particles by cell -> positions
// in cell order (double4 for the exact test, origin-shifted float4 for the pre-filter)
template <int LAYOUT>
static int bin_particles(lj_ctx* ctx, const lj_list_args* a, cudaStream_t st) {
  const int64_t pn = a->pn;
  int rc = lj_scratch_reserve(ctx, pn, st);
  if (rc) return rc;
  unsigned long long* bb = reinterpret_cast<unsigned long long*>(ctx->bbox);
  grid_ext* ge = reinterpret_cast<grid_ext*>(ctx->grid);
  float4* sorted_pos32 = reinterpret_cast<float4*>(ctx->sorted_pos + pn);
  const int* ncell_dev = &ge->ncell1;  // histogram + sentinel
  const int64_t cells = ctx->scratch_cells;
  rc = lj_bbox_launch(ctx, a->q, LAYOUT, pn, a->plane_stride, ctx->totals, st);
  if (rc) return rc;
  k_grid_setup<<<1, 1, 0, st>>>(bb, a->search_len, cells, ge);
  LJ_LAUNCHED(ctx);
  k_zero_u32<<<4 * ctx->sm_count, 256, 0, st>>>(ctx->cell_count, cells + 1, ncell_dev);
  LJ_LAUNCHED(ctx);
  k_cell_assign<LAYOUT><<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(
      a->q, pn, a->plane_stride, ge, ctx->cell_of, ctx->cell_slot, ctx->cell_count);
  LJ_LAUNCHED(ctx);
  // exclusive scan of the cell histogram (length known only on the device)
  const unsigned cell_tiles = (unsigned)blocks_for(cells, kScanTile);
  k_scan_reduce<<<cell_tiles, kScanThreads, 0, st>>>(ctx->cell_count, cells, ncell_dev, ctx->scan_tmp);
  LJ_LAUNCHED(ctx);
  k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, cell_tiles, nullptr);
  LJ_LAUNCHED(ctx);
  k_scan_down<uint32_t><<<cell_tiles, kScanThreads, 0, st>>>(ctx->cell_count, cells, ncell_dev,
                                                              ctx->scan_tmp, ctx->cell_start);
  LJ_LAUNCHED(ctx);
  k_cell_scatter<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, ctx->cell_of, ctx->cell_slot,
                                                                 ctx->cell_start, ctx->sorted_tmp);
  LJ_LAUNCHED(ctx);
  k_cell_order<LAYOUT><<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(
      a->q, pn, a->plane_stride, ge, ctx->cell_of, ctx->cell_start, ctx->cell_count, ctx->sorted_tmp,
      ctx->sorted_pos, sorted_pos32);
  LJ_LAUNCHED(ctx);
  ctx->tl_valid = false;  // the cell-sort scratch and any mirror derived from it are rebuilt from here
  ctx->tl_token = 0;
  ctx->graph_loop = -1;   // ... and a cached CUDA graph may replay the kernel that ran on the old mirror
  return LJ_OK;
}

// *deferred (in: the caller would like to run the FILL pass itself; out: it has to): the COUNT pass
// and the scans are done, sorted_list is still unwritten.  binned: bin_particles() has already run.
template <int LAYOUT>
static int build_list_impl(lj_ctx* ctx, const lj_list_args* a, cudaStream_t st, bool* deferred, bool binned = false) {
  const int64_t pn = a->pn;
  int rc = LJ_OK;
  if (!binned && (rc = bin_particles<LAYOUT>(ctx, a, st))) return rc;
  grid_ext* ge = reinterpret_cast<grid_ext*>(ctx->grid);
  float4* sorted_pos32 = reinterpret_cast<float4*>(ctx->sorted_pos + pn);
  int64_t r0 = a->row_begin, r1 = a->row_end;
  if (r0 == 0 && r1 == 0) r1 = pn;

  const double sl2 = a->search_len * a->search_len;
  if (r0 > 0 || r1 < pn) {  // rows outside the range stay empty
    k_zero_i32_rows<<<4 * ctx->sm_count, 256, 0, st>>>(a->number_of_partners, pn);
    LJ_LAUNCHED(ctx);
  }
  const unsigned row_tiles = (unsigned)blocks_for(pn, kScanTile);
  const uint32_t* nop_u = reinterpret_cast<const uint32_t*>(a->number_of_partners);
  // any list these arrays were mirrored by is stale from here on
  if (ctx->cl_valid && (ctx->cl_id_list == a->sorted_list || ctx->cl_id_nop == a->number_of_partners ||
                        ctx->cl_id_ptr == a->pointer))
    ctx->cl_valid = false;

  // The cluster-organised search is the default engine (one candidate stream per four rows is
  // cheaper than four per-particle searches); LJ_LIST_PER_PARTICLE_SEARCH selects k_search.
  if (!(a->flags & LJ_LIST_PER_PARTICLE_SEARCH) && pn < (1LL << 28) && r1 > r0) {
    // ---------------- cluster search: CSR arrays (+ the cluster pair list on request) -------
    const int64_t nc = (r1 - r0 + 3) / 4;
    if (nc + 1 > ctx->cl_nc_cap) {
      if (ctx->cl_cnt) LJ_CUDA(ctx, cudaFreeAsync(ctx->cl_cnt, st));
      if (ctx->cl_ptr) LJ_CUDA(ctx, cudaFreeAsync(ctx->cl_ptr, st));
      ctx->cl_cnt = nullptr; ctx->cl_ptr = nullptr;
      LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cl_cnt, sizeof(uint32_t) * (nc + 1), ctx->pool, st));
      LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cl_ptr, sizeof(long long) * (nc + 1), ctx->pool, st));
      ctx->cl_nc_cap = nc + 1;
      ctx->graph_loop = -1;  // a cached CUDA graph may hold the old pointers: force a recapture
    }
    const bool emit = !a->half && (a->flags & LJ_LIST_CLUSTERS);
    LJ_CUDA(ctx, cudaMemsetAsync(ctx->cl_cnt + nc, 0, sizeof(uint32_t), st));
    const unsigned cblocks = (unsigned)blocks_for(nc * kSearchLanes, 256);
    k_search_cluster<false, false, LAYOUT><<<cblocks, 256, 0, st>>>(
        a->q, a->plane_stride, pn, ge, ctx->cell_start, ctx->sorted_pos, sorted_pos32, sl2, a->half, r0,
        r1, a->number_of_partners, nullptr, nullptr, 0, emit ? ctx->cl_cnt : nullptr, nullptr, nullptr, 0,
        ctx->totals);
    LJ_LAUNCHED(ctx);
    k_scan_reduce<<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp);
    LJ_LAUNCHED(ctx);
    k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, row_tiles, &ctx->totals->total);
    LJ_LAUNCHED(ctx);
    if (a->pointer64)
      k_scan_down<long long><<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp,
                                                                  reinterpret_cast<long long*>(a->pointer));
    else
      k_scan_down<uint32_t><<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp,
                                                                 reinterpret_cast<uint32_t*>(a->pointer));
    LJ_LAUNCHED(ctx);
    if (emit) {
      const unsigned ctiles = (unsigned)blocks_for(nc + 1, kScanTile);
      k_scan_reduce<<<ctiles, kScanThreads, 0, st>>>(ctx->cl_cnt, nc + 1, nullptr, ctx->scan_tmp);
      LJ_LAUNCHED(ctx);
      k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, ctiles, &ctx->totals->cl_total);
      LJ_LAUNCHED(ctx);
      k_scan_down<long long><<<ctiles, kScanThreads, 0, st>>>(ctx->cl_cnt, nc + 1, nullptr, ctx->scan_tmp,
                                                               ctx->cl_ptr);
      LJ_LAUNCHED(ctx);
    }
    k_finish_totals<<<1, 1, 0, st>>>(ctx->totals, a->capacity, a->pointer64);
    LJ_LAUNCHED(ctx);
    ctx->last_capacity = a->capacity;
    int64_t need = 0;
    if (emit) {
      // one small read-back per build: sizes the library-owned cluster list and lets an
      // over-capacity build skip the fill pass altogether
      LJ_CUDA(ctx, cudaMemcpyAsync(ctx->totals_host, ctx->totals, sizeof(lj_list_totals),
                                   cudaMemcpyDeviceToHost, st));
      LJ_CUDA(ctx, cudaStreamSynchronize(st));
      if (ctx->totals_host->overflow) return LJ_OK;  // reported by lj_list_result
      need = (int64_t)ctx->totals_host->cl_total;
    }
    if (emit && need > ctx->cl_cap) {
      if (ctx->cl_list) LJ_CUDA(ctx, cudaFreeAsync(ctx->cl_list, st));
      ctx->cl_list = nullptr;
      ctx->cl_cap = need + need / 32 + 4096;
      ctx->graph_loop = -1;  // see above
      LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->cl_list, sizeof(uint32_t) * ctx->cl_cap, ctx->pool, st));
    }
    if (*deferred && !emit) return LJ_OK;  // the cell-tile fill pass writes sorted_list as well
    *deferred = false;
    rc = cluster_fill<LAYOUT>(ctx, a, st, emit);
    if (rc) return rc;
    if (emit) {
      ctx->cl_valid = true;
      ctx->cl_id_list = a->sorted_list; ctx->cl_id_nop = a->number_of_partners; ctx->cl_id_ptr = a->pointer;
      ctx->cl_pn = pn; ctx->cl_r0 = r0; ctx->cl_r1 = r1; ctx->cl_entries = need;
    }
    return LJ_OK;
  }

  *deferred = false;
  const unsigned search_blocks = (unsigned)blocks_for(pn * kSearchLanes, 256);
  k_search<false, false><<<search_blocks, 256, 0, st>>>(
      pn, ge, ctx->cell_of, ctx->cell_start, ctx->sorted_pos, sorted_pos32, sl2,
      a->half, r0, r1, a->number_of_partners, nullptr, nullptr, 0, ctx->totals);
  LJ_LAUNCHED(ctx);
  // pointer[] = exclusive scan of number_of_partners, carried in 64 bits
  k_scan_reduce<<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp);
  LJ_LAUNCHED(ctx);
  k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, row_tiles, &ctx->totals->total);
  LJ_LAUNCHED(ctx);
  if (a->pointer64)
    k_scan_down<long long><<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp,
                                                                reinterpret_cast<long long*>(a->pointer));
  else
    k_scan_down<uint32_t><<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp,
                                                               reinterpret_cast<uint32_t*>(a->pointer));
  LJ_LAUNCHED(ctx);
  k_finish_totals<<<1, 1, 0, st>>>(ctx->totals, a->capacity, a->pointer64);
  LJ_LAUNCHED(ctx);
  if (a->pointer64)
    k_search<true, true><<<search_blocks, 256, 0, st>>>(
        pn, ge, ctx->cell_of, ctx->cell_start, ctx->sorted_pos, sorted_pos32, sl2,
        a->half, r0, r1, a->number_of_partners, a->pointer, a->sorted_list, a->capacity, ctx->totals);
  else
    k_search<true, false><<<search_blocks, 256, 0, st>>>(
        pn, ge, ctx->cell_of, ctx->cell_start, ctx->sorted_pos, sorted_pos32, sl2,
        a->half, r0, r1, a->number_of_partners, a->pointer, a->sorted_list, a->capacity, ctx->totals);
  LJ_LAUNCHED(ctx);
  ctx->last_capacity = a->capacity;
  return LJ_OK;
}


template <int LAYOUT>
static int cluster_fill(lj_ctx* ctx, const lj_list_args* a, cudaStream_t st, bool emit) {
  const int64_t pn = a->pn;
  int64_t r0 = a->row_begin, r1 = a->row_end;
  if (r0 == 0 && r1 == 0) r1 = pn;
  grid_ext* ge = reinterpret_cast<grid_ext*>(ctx->grid);
  float4* sorted_pos32 = reinterpret_cast<float4*>(ctx->sorted_pos + pn);
  const double sl2 = a->search_len * a->search_len;
  const int64_t nc = (r1 - r0 + 3) / 4;
  const unsigned cblocks = (unsigned)blocks_for(nc * kSearchLanes, 256);
  if (a->pointer64)
    k_search_cluster<true, true, LAYOUT><<<cblocks, 256, 0, st>>>(
        a->q, a->plane_stride, pn, ge, ctx->cell_start, ctx->sorted_pos, sorted_pos32, sl2, a->half, r0,
        r1, a->number_of_partners, a->pointer, a->sorted_list, a->capacity,
        emit ? ctx->cl_cnt : nullptr, ctx->cl_ptr, ctx->cl_list, ctx->cl_cap, ctx->totals);
  else
    k_search_cluster<true, false, LAYOUT><<<cblocks, 256, 0, st>>>(
        a->q, a->plane_stride, pn, ge, ctx->cell_start, ctx->sorted_pos, sorted_pos32, sl2, a->half, r0,
        r1, a->number_of_partners, a->pointer, a->sorted_list, a->capacity,
        emit ? ctx->cl_cnt : nullptr, ctx->cl_ptr, ctx->cl_list, ctx->cl_cap, ctx->totals);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

// ------------------------------------------------------------------ cell-tile mirror: host
// Runs right after build_list_impl on the same stream: the cell-sort scratch (cell_of, cell_start,
// sorted_pos*) still describes a->q.  Two small read-backs: the tile count (sizes the per-tile
// table and the list) and the per-tile maxima (size the force kernel's shared memory).
static int tile_alloc(lj_ctx* ctx, void** ptr, size_t bytes, cudaStream_t st) {
  if (*ptr) LJ_CUDA(ctx, cudaFreeAsync(*ptr, st));
  *ptr = nullptr;
  LJ_CUDA(ctx, cudaMallocAsync(ptr, bytes, ctx->pool, st));
  ctx->graph_loop = -1;  // a cached CUDA graph may hold the old pointer
  return LJ_OK;
}

static int build_tile_mirror(lj_ctx* ctx, const lj_list_args* a, cudaStream_t st, bool fill_public,
                             bool* filled_public) {
  *filled_public = false;
  const int64_t pn = a->pn;
  int64_t r0 = a->row_begin, r1 = a->row_end;
  if (r0 == 0 && r1 == 0) r1 = pn;
  grid_ext* ge = reinterpret_cast<grid_ext*>(ctx->grid);
  float4* sorted_pos32 = reinterpret_cast<float4*>(ctx->sorted_pos + pn);
  int rc;
  if (!ctx->tl_geom) {
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->tl_geom, sizeof(lj_tile_geom), ctx->pool, st));
    LJ_CUDA(ctx, cudaHostAlloc((void**)&ctx->tl_geom_host, sizeof(lj_tile_geom), cudaHostAllocDefault));
  }
  if (pn > ctx->tl_pn_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_order, sizeof(int32_t) * pn, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cnt, sizeof(int32_t) * pn, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_units, sizeof(uint32_t) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_off, sizeof(uint32_t) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_meta, sizeof(int4) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_qs, 24 * (size_t)(pn + 2), st))) return rc;
    LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_qs, 0, 24 * (size_t)(pn + 2), st));
    ctx->tl_qz = ctx->tl_qs + 2 * (size_t)(pn + 2);
    ctx->tl_pn_cap = pn;
  }
  const int rows_env = lj_diag_int("LJ_TILE_ROWS");
  const int target_rows = rows_env > 0 ? rows_env : (a->flags & LJ_LIST_TILES_WIDE) ? LJ_TILE_ROWS_WIDE : LJ_TILE_ROWS_STD;
  k_tile_prepare<<<1, 1, 0, st>>>(ge, pn, target_rows, ctx->tl_geom);
  LJ_LAUNCHED(ctx);
  k_tile_rows<<<(unsigned)blocks_for(pn + 1, 256), 256, 0, st>>>(pn, sorted_pos32, a->number_of_partners,
                                                                  ctx->tl_order, ctx->tl_cnt, ctx->tl_units);
  LJ_LAUNCHED(ctx);
  const unsigned tiles = (unsigned)blocks_for(pn + 1, kScanTile);
  k_scan_reduce<<<tiles, kScanThreads, 0, st>>>(ctx->tl_units, pn + 1, nullptr, ctx->scan_tmp);
  LJ_LAUNCHED(ctx);
  k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, tiles, &ctx->tl_geom->total_units);
  LJ_LAUNCHED(ctx);
  k_scan_down<uint32_t><<<tiles, kScanThreads, 0, st>>>(ctx->tl_units, pn + 1, nullptr, ctx->scan_tmp,
                                                         ctx->tl_off);
  LJ_LAUNCHED(ctx);
  k_tile_meta<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, ctx->tl_order, ctx->tl_cnt, ctx->tl_off, ctx->tl_meta);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->totals_host, ctx->totals, sizeof(lj_list_totals), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  if (ctx->totals_host->overflow) { *filled_public = true; return LJ_OK; }  // reported by lj_list_result; nothing to fill
  lj_tile_geom g = *ctx->tl_geom_host;
  if (g.total_units >= 0xffffffffull) return LJ_OK;  // 32-bit unit offsets: no mirror, per-row kernels serve
  const int64_t ncell1 = (int64_t)g.nx * g.ny * g.nz + 1;
  if (ncell1 > ctx->tl_cells_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cell_start, sizeof(uint32_t) * ncell1, st))) return rc;
    ctx->tl_cells_cap = ncell1;
  }
  if (g.ntiles > ctx->tl_tab_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_tab, sizeof(uint2) * kTileYTab * (size_t)g.ntiles, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_ttab, sizeof(uint4) * kTileTTab * (size_t)g.ntiles, st))) return rc;
    ctx->tl_tab_cap = g.ntiles;
  }
  if ((int64_t)g.total_units + 2 > ctx->tl_list_cap) {
    const int64_t cap = (int64_t)g.total_units + (int64_t)g.total_units / 32 + 1024;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_list, 16 * (size_t)cap, st))) return rc;
    ctx->tl_list_cap = cap;
  }
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_cell_start, ctx->cell_start, sizeof(uint32_t) * ncell1,
                               cudaMemcpyDeviceToDevice, st));
  const int ncols_all = g.ntx * g.nz;
  if (ncols_all > ctx->tl_cols_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cols, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    ctx->tl_cols_cap = ncols_all;
  }
  if (ncols_all > ctx->tl_sel_cap) {  // column selection of part launches: z-layer flags, compacted columns
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_zflag, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cols_sel, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    ctx->tl_sel_cap = ncols_all;
  }
  LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_zflag, 0, sizeof(int32_t) * (size_t)ncols_all, st));
  if (r0 > 0 || r1 < pn) {
    k_tile_zflag<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, r0, r1, ctx->cell_of, ctx->tl_geom, ctx->tl_zflag);
    LJ_LAUNCHED(ctx);
  }
  LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_cols, 0, sizeof(int32_t) * (size_t)ncols_all, st));
  k_tile_table<<<(unsigned)blocks_for(g.ntiles, 128), 128, 0, st>>>(ctx->tl_cell_start, ctx->tl_off,
                                                                     ctx->tl_geom, ctx->tl_tab, ctx->tl_ttab, ctx->tl_cols, 2);
  LJ_LAUNCHED(ctx);
  k_tile_cols<<<1, 32, 0, st>>>(ncols_all, ctx->tl_cols, ctx->tl_geom);
  LJ_LAUNCHED(ctx);
  const unsigned fblocks = (unsigned)blocks_for(pn * kSearchLanes, 256);
#define LJ_TILE_FILL(PUB, P64)                                                                        \
  k_tile_fill<PUB, P64><<<fblocks, 256, 0, st>>>(                                                     \
      pn, ge, ctx->tl_geom, ctx->cell_of, ctx->tl_cell_start, ctx->sorted_pos, sorted_pos32,          \
      a->search_len * a->search_len, ctx->tl_cnt, ctx->tl_off, ctx->tl_tab, ctx->tl_list, ctx->totals, \
      a->pointer, a->sorted_list, a->capacity, lj_diag_set("LJ_TILE_FAKE") ? 1 : 0)
  if (!fill_public) LJ_TILE_FILL(false, false);
  else if (a->pointer64) LJ_TILE_FILL(true, true);
  else LJ_TILE_FILL(true, false);
#undef LJ_TILE_FILL
  LJ_LAUNCHED(ctx);
  *filled_public = fill_public;
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->totals_host, ctx->totals, sizeof(lj_list_totals), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  if (ctx->totals_host->overflow & 8)
    return lj_set_error(ctx, LJ_ERR_INVALID_LIST, "lj_build_list", "cell-tile mirror disagrees with the CSR list");
  if (ctx->totals_host->overflow) return LJ_OK;  // capacity problems are reported by lj_list_result
  g = *ctx->tl_geom_host;
  if (lj_diag_set("LJ_TILE_DEBUG"))
    fprintf(stderr, "[lj] cell-tile mirror: grid %dx%dx%d, %d cells per tile, %d tiles, max rows %d, y-row %d, "
            "list units %d, y slot %zu B, list slot %zu B, %llu units total\n", g.nx, g.ny, g.nz, g.tc,
            g.ntiles, g.max_rows, g.max_yrow, g.max_units, lj_celltile_yslot_bytes(g),
            lj_celltile_lslot_bytes(g), g.total_units);
  // too dense for 16-bit indices or for the shared-memory rings: no mirror, the per-row kernels serve
  if (5 * lj_celltile_cap_y(g) >= 65536 ||
      kTileMinYSlots * lj_celltile_yslot_bytes(g) + kTileMinLSlots * lj_celltile_lslot_bytes(g) > kTileSmemBudget)
    return LJ_OK;
  ctx->tl_g = g;
  ctx->tl_valid = true;
  ctx->tl_token = ctx->tl_token_next++;
  ctx->tl_outside = 0;
  ctx->tl_id_list = a->sorted_list; ctx->tl_id_nop = a->number_of_partners; ctx->tl_id_ptr = a->pointer;
  ctx->tl_pn = pn; ctx->tl_r0 = r0; ctx->tl_r1 = r1;
  ctx->graph_loop = -1;  // the geometry is baked into the launch parameters
  return LJ_OK;
}

// ------------------------------------------------------------------ tile engine: host
// lj_build_list with LJ_LIST_TILES on a full list: binning -> tile geometry -> k_tile_count (the ONE
// search, masks) -> scans (pointer[], mirror offsets) -> tables -> k_tile_replay (both lists).
// *done = false: the engine does not apply to this system (cells too crowded for the 64-bit window
// masks, a region that does not fit in shared memory, mask scratch too large, > 2^32 mirror units);
// the binning has been done, the caller continues with the round-1 engine.
template <int LAYOUT>
static int build_list_tiles(lj_ctx* ctx, const lj_list_args* a_in, cudaStream_t st, bool* done) {
  *done = false;
  lj_list_args a_local = *a_in;
  lj_list_args* a = &a_local;
  const bool self_alloc = (a->flags & LJ_LIST_ALLOC_INTERNAL) && !a->sorted_list && a->capacity == 0;
  ctx->alloc_list = nullptr; ctx->alloc_capacity = 0;
  const int64_t pn = a->pn;
  int64_t r0 = a->row_begin, r1 = a->row_end;
  if (r0 == 0 && r1 == 0) r1 = pn;
  int rc = bin_particles<LAYOUT>(ctx, a, st);
  if (rc) return rc;
  const size_t mask_bytes = sizeof(unsigned long long) * 25 * (size_t)pn;
  if (mask_bytes > ((size_t)16 << 30)) {
    // A big system (config 5 on one GPU: 26 GB of masks beside a 70 GB list): only when the device has the
    // room -- what the driver reports free plus what this context's pool holds cached -- for the masks, the
    // mirror (2 B per entry + padding, unless one of that size exists) and the per-row arrays, with 8 GiB to
    // spare; otherwise the round-1 engine builds the list without mask scratch.
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return LJ_OK; }
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      free_b += (size_t)(reserved - used);
    else
      cudaGetLastError();
    const size_t entries = a->capacity > 0 ? (size_t)a->capacity : (size_t)pn * 160;
    size_t need = mask_bytes + ((size_t)8 << 30);
    if (16 * (size_t)ctx->tl_list_cap < entries * 2) need += entries * 2 + entries / 4;
    if (pn > ctx->tl_pn_cap) need += (size_t)pn * 64;
    if (free_b < need) return LJ_OK;
  }
  grid_ext* ge = reinterpret_cast<grid_ext*>(ctx->grid);
  float4* sorted_pos32 = reinterpret_cast<float4*>(ctx->sorted_pos + pn);
  if (ctx->cl_valid && (ctx->cl_id_list == a->sorted_list || ctx->cl_id_nop == a->number_of_partners ||
                        ctx->cl_id_ptr == a->pointer))
    ctx->cl_valid = false;
  if (!ctx->tl_geom) {
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->tl_geom, sizeof(lj_tile_geom), ctx->pool, st));
    LJ_CUDA(ctx, cudaHostAlloc((void**)&ctx->tl_geom_host, sizeof(lj_tile_geom), cudaHostAllocDefault));
  }
  if (pn > ctx->tl_pn_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_order, sizeof(int32_t) * pn, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cnt, sizeof(int32_t) * pn, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_units, sizeof(uint32_t) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_off, sizeof(uint32_t) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_meta, sizeof(int4) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_qs, 24 * (size_t)(pn + 2), st))) return rc;
    LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_qs, 0, 24 * (size_t)(pn + 2), st));
    ctx->tl_qz = ctx->tl_qs + 2 * (size_t)(pn + 2);
    ctx->tl_pn_cap = pn;
  }
  const int rows_env = lj_diag_int("LJ_TILE_ROWS");
  const int target_rows = rows_env > 0 ? rows_env : (a->flags & LJ_LIST_TILES_WIDE) ? LJ_TILE_ROWS_WIDE : LJ_TILE_ROWS_STD;
  k_tile_prepare<<<1, 1, 0, st>>>(ge, pn, target_rows, ctx->tl_geom);
  LJ_LAUNCHED(ctx);
  // read-back 1: the grid and the tile count (size the tables and the launches)
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  lj_tile_geom g = *ctx->tl_geom_host;
  if (g.ntiles <= 0) return LJ_OK;
  const int64_t ncell1 = (int64_t)g.nx * g.ny * g.nz + 1;
  if (ncell1 > ctx->tl_cells_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cell_start, sizeof(uint32_t) * ncell1, st))) return rc;
    ctx->tl_cells_cap = ncell1;
  }
  if (g.ntiles > ctx->tl_tab_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_tab, sizeof(uint2) * kTileYTab * (size_t)g.ntiles, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_ttab, sizeof(uint4) * kTileTTab * (size_t)g.ntiles, st))) return rc;
    ctx->tl_tab_cap = g.ntiles;
  }
  const int ncols_all = g.ntx * g.nz;
  if (ncols_all > ctx->tl_cols_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cols, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    ctx->tl_cols_cap = ncols_all;
  }
  if (ncols_all > ctx->tl_sel_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_zflag, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cols_sel, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    ctx->tl_sel_cap = ncols_all;
  }
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_cell_start, ctx->cell_start, sizeof(uint32_t) * ncell1,
                               cudaMemcpyDeviceToDevice, st));
  k_tile_table<<<(unsigned)blocks_for(g.ntiles, 128), 128, 0, st>>>(ctx->tl_cell_start, ctx->tl_off, ctx->tl_geom,
                                                                     ctx->tl_tab, ctx->tl_ttab, ctx->tl_cols, 0);
  LJ_LAUNCHED(ctx);
  // read-back 2: the longest y-row (sizes the shared memory of the count kernel)
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  g = *ctx->tl_geom_host;
  const int cap_y = lj_celltile_cap_y(g);
  const int ncx1 = g.tc + 5;  // x-cells of a region + 1
  const size_t smem_tab = 25 * sizeof(te_pencil) + sizeof(int) * 25 * (size_t)ncx1;
  const size_t smem_count = (size_t)(5 * cap_y + 64) * sizeof(float4) + smem_tab + kTcRowChunk * (sizeof(int) + 1);
  if (5 * cap_y >= 65536 || smem_count > (size_t)200 * 1024) return LJ_OK;
  unsigned long long* masks = nullptr;
  LJ_CUDA(ctx, cudaMallocAsync((void**)&masks, mask_bytes, ctx->pool, st));
  LJ_FUNC_SMEM(ctx, k_tile_count2, smem_count);
  const double sl2 = a->search_len * a->search_len;
  k_tile_count2<<<(unsigned)g.ntiles, kTcThreads, smem_count, st>>>(
      pn, r0, r1, ge, ctx->tl_geom, ctx->tl_cell_start, ctx->tl_tab, ctx->sorted_pos, sorted_pos32, sl2, ncx1,
      a->number_of_partners, ctx->tl_order, ctx->tl_cnt, ctx->tl_units, masks, ctx->totals);
  LJ_LAUNCHED(ctx);
  // pointer[] = exclusive scan of number_of_partners (64-bit carry), mirror offsets = scan of the padded units
  const unsigned row_tiles = (unsigned)blocks_for(pn, kScanTile);
  const uint32_t* nop_u = reinterpret_cast<const uint32_t*>(a->number_of_partners);
  k_scan_reduce<<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp);
  LJ_LAUNCHED(ctx);
  k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, row_tiles, &ctx->totals->total);
  LJ_LAUNCHED(ctx);
  if (a->pointer64)
    k_scan_down<long long><<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp,
                                                                reinterpret_cast<long long*>(a->pointer));
  else
    k_scan_down<uint32_t><<<row_tiles, kScanThreads, 0, st>>>(nop_u, pn, nullptr, ctx->scan_tmp,
                                                               reinterpret_cast<uint32_t*>(a->pointer));
  LJ_LAUNCHED(ctx);
  k_finish_totals<<<1, 1, 0, st>>>(ctx->totals, self_alloc ? (int64_t)0x7fffffffffffffffLL : a->capacity, a->pointer64);
  LJ_LAUNCHED(ctx);
  ctx->last_capacity = a->capacity;
  const unsigned utiles = (unsigned)blocks_for(pn + 1, kScanTile);
  k_scan_reduce<<<utiles, kScanThreads, 0, st>>>(ctx->tl_units, pn + 1, nullptr, ctx->scan_tmp);
  LJ_LAUNCHED(ctx);
  k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, utiles, &ctx->tl_geom->total_units);
  LJ_LAUNCHED(ctx);
  k_scan_down<uint32_t><<<utiles, kScanThreads, 0, st>>>(ctx->tl_units, pn + 1, nullptr, ctx->scan_tmp, ctx->tl_off);
  LJ_LAUNCHED(ctx);
  k_tile_meta<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, ctx->tl_order, ctx->tl_cnt, ctx->tl_off, ctx->tl_meta);
  LJ_LAUNCHED(ctx);
  // tile table, active columns, z-layer flags: need the offsets, not the list itself
  LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_zflag, 0, sizeof(int32_t) * (size_t)ncols_all, st));
  if (r0 > 0 || r1 < pn) {
    k_tile_zflag<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, r0, r1, ctx->cell_of, ctx->tl_geom, ctx->tl_zflag);
    LJ_LAUNCHED(ctx);
  }
  LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_cols, 0, sizeof(int32_t) * (size_t)ncols_all, st));
  k_tile_table<<<(unsigned)blocks_for(g.ntiles, 128), 128, 0, st>>>(ctx->tl_cell_start, ctx->tl_off, ctx->tl_geom,
                                                                     ctx->tl_tab, ctx->tl_ttab, ctx->tl_cols, 1);
  LJ_LAUNCHED(ctx);
  k_tile_cols<<<1, 32, 0, st>>>(ncols_all, ctx->tl_cols, ctx->tl_geom);
  LJ_LAUNCHED(ctx);
  // read-back 3: totals (capacity / offset overflow, crowded cells) and the size of the mirror
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->totals_host, ctx->totals, sizeof(lj_list_totals), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  g = *ctx->tl_geom_host;
  const int ovf = ctx->totals_host->overflow;
  if ((ovf & 16) || g.total_units >= 0xffffffffull) {  // not for this engine: the round-1 engine redoes the counts
    LJ_CUDA(ctx, cudaFreeAsync(masks, st));
    LJ_CUDA(ctx, cudaMemsetAsync(ctx->totals, 0, sizeof(lj_list_totals), st));  // its flags start clean
    return LJ_OK;
  }
  *done = true;
  if (ovf) {  // capacity or 32-bit offset overflow: reported by lj_list_result, nothing to fill
    LJ_CUDA(ctx, cudaFreeAsync(masks, st));
    return LJ_OK;
  }
  if ((int64_t)g.total_units + 2 > ctx->tl_list_cap) {
    const int64_t cap = (int64_t)g.total_units + (int64_t)g.total_units / 32 + 1024;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_list, 16 * (size_t)cap, st))) return rc;
    ctx->tl_list_cap = cap;
  }
  if (self_alloc) {  // lj_measure: the total is known now -- no separate sizing build
    const int64_t total = (int64_t)ctx->totals_host->total;
    ctx->alloc_capacity = total + total / 64 + 1024;  // headroom for later rebuilds
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->alloc_list, sizeof(int32_t) * (size_t)ctx->alloc_capacity, ctx->pool, st));
    a->sorted_list = ctx->alloc_list; a->capacity = ctx->alloc_capacity;
    ctx->last_capacity = a->capacity;
  }
  // staged entries of a tile (6 B each) + tables + per-row arrays (see k_tile_replay2)
  int rowcap = (ctx->totals_host->max_np + 7) & ~7;  // (read-back 3) a multiple of 8: the 4-byte plane stays aligned
  rowcap = rowcap < 64 ? 64 : rowcap > kTeRowCap ? kTeRowCap : rowcap;
  const size_t smem_replay = smem_tab + (size_t)(kTeThreads / kTeLanes) * te_group_stride(rowcap);
  LJ_FUNC_SMEM(ctx, k_tile_replay<true>, smem_replay);
  LJ_FUNC_SMEM(ctx, k_tile_replay<false>, smem_replay);
  if (a->pointer64)
    k_tile_replay<true><<<(unsigned)g.ntiles, kTeThreads, smem_replay, st>>>(
        pn, ctx->tl_geom, ctx->tl_cell_start, ctx->tl_tab, ncx1, ctx->tl_order, ctx->tl_cnt, ctx->tl_off, masks, ctx->tl_list,
        a->pointer, a->sorted_list, a->capacity, rowcap, ctx->totals);
  else
    k_tile_replay<false><<<(unsigned)g.ntiles, kTeThreads, smem_replay, st>>>(
        pn, ctx->tl_geom, ctx->tl_cell_start, ctx->tl_tab, ncx1, ctx->tl_order, ctx->tl_cnt, ctx->tl_off, masks, ctx->tl_list,
        a->pointer, a->sorted_list, a->capacity, rowcap, ctx->totals);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaFreeAsync(masks, st));
  // read-back 4 is not needed: the geometry is final since read-back 3 (k_tile_cols included); the
  // replay can only raise bit 8 (masks disagree with the counts), which lj_list_result reports
  if (lj_diag_set("LJ_TILE_DEBUG"))
    fprintf(stderr, "[lj] cell-tile mirror (tile engine): grid %dx%dx%d, %d cells per tile, %d tiles, max rows %d, y-row %d, "
            "list units %d, y slot %zu B, list slot %zu B, %llu units total\n", g.nx, g.ny, g.nz, g.tc,
            g.ntiles, g.max_rows, g.max_yrow, g.max_units, lj_celltile_yslot_bytes(g),
            lj_celltile_lslot_bytes(g), g.total_units);
  // too dense for the shared-memory rings of the force kernel: the lists are complete, there is no mirror
  if (kTileMinYSlots * lj_celltile_yslot_bytes(g) + kTileMinLSlots * lj_celltile_lslot_bytes(g) > kTileSmemBudget)
    return LJ_OK;
  ctx->tl_g = g;
  ctx->tl_valid = true;
  ctx->tl_token = ctx->tl_token_next++;
  ctx->tl_outside = 0;
  ctx->tl_id_list = a->sorted_list; ctx->tl_id_nop = a->number_of_partners; ctx->tl_id_ptr = a->pointer;
  ctx->tl_pn = pn; ctx->tl_r0 = r0; ctx->tl_r1 = r1;
  ctx->graph_loop = -1;  // the geometry is baked into the launch parameters
  return LJ_OK;
}

template <int LAYOUT>
static int list_mirror_impl(lj_ctx* ctx, const lj_list_args* a, int64_t* rows_outside_out, cudaStream_t st) {
  const int64_t pn = a->pn;
  int rc = bin_particles<LAYOUT>(ctx, a, st);
  if (rc) return rc;
  grid_ext* ge = reinterpret_cast<grid_ext*>(ctx->grid);
  float4* sorted_pos32 = reinterpret_cast<float4*>(ctx->sorted_pos + pn);
  if (!ctx->tl_geom) {
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->tl_geom, sizeof(lj_tile_geom), ctx->pool, st));
    LJ_CUDA(ctx, cudaHostAlloc((void**)&ctx->tl_geom_host, sizeof(lj_tile_geom), cudaHostAllocDefault));
  }
  if (pn > ctx->tl_pn_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_order, sizeof(int32_t) * pn, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cnt, sizeof(int32_t) * pn, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_units, sizeof(uint32_t) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_off, sizeof(uint32_t) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_meta, sizeof(int4) * (pn + 1), st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_qs, 24 * (size_t)(pn + 2), st))) return rc;
    LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_qs, 0, 24 * (size_t)(pn + 2), st));
    ctx->tl_qz = ctx->tl_qs + 2 * (size_t)(pn + 2);
    ctx->tl_pn_cap = pn;
  }
  if (pn > ctx->tl_rowflag_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_rowflag, (size_t)pn, st))) return rc;
    ctx->tl_rowflag_cap = pn;
  }
  if (pn > ctx->tl_slot_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_slot_of, sizeof(int32_t) * (size_t)pn, st))) return rc;
    ctx->tl_slot_cap = pn;
  }
  const int target_rows = (a->flags & LJ_LIST_TILES_WIDE) ? LJ_TILE_ROWS_WIDE : LJ_TILE_ROWS_STD;
  k_tile_prepare<<<1, 1, 0, st>>>(ge, pn, target_rows, ctx->tl_geom);
  LJ_LAUNCHED(ctx);
  // rows in cell order with the caller's counts, their padded lengths and offsets
  k_tile_rows<<<(unsigned)blocks_for(pn + 1, 256), 256, 0, st>>>(pn, sorted_pos32, a->number_of_partners,
                                                                  ctx->tl_order, ctx->tl_cnt, ctx->tl_units);
  LJ_LAUNCHED(ctx);
  const unsigned utiles = (unsigned)blocks_for(pn + 1, kScanTile);
  k_scan_reduce<<<utiles, kScanThreads, 0, st>>>(ctx->tl_units, pn + 1, nullptr, ctx->scan_tmp);
  LJ_LAUNCHED(ctx);
  k_scan_spine<<<1, kScanThreads, 0, st>>>(ctx->scan_tmp, utiles, &ctx->tl_geom->total_units);
  LJ_LAUNCHED(ctx);
  k_scan_down<uint32_t><<<utiles, kScanThreads, 0, st>>>(ctx->tl_units, pn + 1, nullptr, ctx->scan_tmp, ctx->tl_off);
  LJ_LAUNCHED(ctx);
  k_tile_meta<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, ctx->tl_order, ctx->tl_cnt, ctx->tl_off, ctx->tl_meta);
  LJ_LAUNCHED(ctx);
  k_slot_of<<<(unsigned)blocks_for(pn, 256), 256, 0, st>>>(pn, ctx->tl_order, ctx->tl_slot_of);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  lj_tile_geom g = *ctx->tl_geom_host;
  LJ_REQUIRE(ctx, g.ntiles > 0 && g.total_units < 0xffffffffull, "lj_list_mirror: list too large for 32-bit mirror offsets");
  const int64_t ncell1 = (int64_t)g.nx * g.ny * g.nz + 1;
  if (ncell1 > ctx->tl_cells_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cell_start, sizeof(uint32_t) * ncell1, st))) return rc;
    ctx->tl_cells_cap = ncell1;
  }
  if (g.ntiles > ctx->tl_tab_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_tab, sizeof(uint2) * kTileYTab * (size_t)g.ntiles, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_ttab, sizeof(uint4) * kTileTTab * (size_t)g.ntiles, st))) return rc;
    ctx->tl_tab_cap = g.ntiles;
  }
  const int ncols_all = g.ntx * g.nz;
  if (ncols_all > ctx->tl_cols_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cols, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    ctx->tl_cols_cap = ncols_all;
  }
  if (ncols_all > ctx->tl_sel_cap) {
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_zflag, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_cols_sel, sizeof(int32_t) * (size_t)ncols_all, st))) return rc;
    ctx->tl_sel_cap = ncols_all;
  }
  if ((int64_t)g.total_units + 2 > ctx->tl_list_cap) {
    const int64_t cap = (int64_t)g.total_units + (int64_t)g.total_units / 32 + 1024;
    if ((rc = tile_alloc(ctx, (void**)&ctx->tl_list, 16 * (size_t)cap, st))) return rc;
    ctx->tl_list_cap = cap;
  }
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_cell_start, ctx->cell_start, sizeof(uint32_t) * ncell1,
                               cudaMemcpyDeviceToDevice, st));
  LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_zflag, 0, sizeof(int32_t) * (size_t)ncols_all, st));
  LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_cols, 0, sizeof(int32_t) * (size_t)ncols_all, st));
  k_tile_table<<<(unsigned)blocks_for(g.ntiles, 128), 128, 0, st>>>(ctx->tl_cell_start, ctx->tl_off, ctx->tl_geom,
                                                                     ctx->tl_tab, ctx->tl_ttab, ctx->tl_cols, 2);
  LJ_LAUNCHED(ctx);
  k_tile_cols<<<1, 32, 0, st>>>(ncols_all, ctx->tl_cols, ctx->tl_geom);
  LJ_LAUNCHED(ctx);
  const int ncx1 = g.tc + 5;
  const size_t smem_tab = 25 * sizeof(te_pencil) + sizeof(int) * 25 * (size_t)ncx1;
  LJ_REQUIRE(ctx, smem_tab <= (size_t)200 * 1024, "lj_list_mirror: tile too wide");
  LJ_FUNC_SMEM(ctx, k_tile_translate<true>, smem_tab);
  LJ_FUNC_SMEM(ctx, k_tile_translate<false>, smem_tab);
  if (a->pointer64)
    k_tile_translate<true><<<(unsigned)g.ntiles, kTeThreads, smem_tab, st>>>(
        pn, ctx->tl_geom, ctx->tl_cell_start, ctx->tl_tab, ncx1, ctx->cell_of, ctx->tl_slot_of, ctx->tl_order,
        ctx->tl_cnt, ctx->tl_off, ctx->tl_meta, ctx->tl_list, a->pointer, a->sorted_list, ctx->tl_rowflag, ctx->totals);
  else
    k_tile_translate<false><<<(unsigned)g.ntiles, kTeThreads, smem_tab, st>>>(
        pn, ctx->tl_geom, ctx->tl_cell_start, ctx->tl_tab, ncx1, ctx->cell_of, ctx->tl_slot_of, ctx->tl_order,
        ctx->tl_cnt, ctx->tl_off, ctx->tl_meta, ctx->tl_list, a->pointer, a->sorted_list, ctx->tl_rowflag, ctx->totals);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->tl_geom_host, ctx->tl_geom, sizeof(lj_tile_geom), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->totals_host, ctx->totals, sizeof(lj_list_totals), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  g = *ctx->tl_geom_host;
  const int64_t outside = (int64_t)ctx->totals_host->cl_total;
  if (rows_outside_out) *rows_outside_out = outside;
  LJ_REQUIRE(ctx, 5 * lj_celltile_cap_y(g) < 65536 &&
                 kTileMinYSlots * lj_celltile_yslot_bytes(g) + kTileMinLSlots * lj_celltile_lslot_bytes(g) <= kTileSmemBudget,
             "lj_list_mirror: a tile's neighbourhood does not fit in shared memory (system too dense)");
  ctx->tl_g = g;
  ctx->tl_valid = true;
  ctx->tl_token = ctx->tl_token_next++;
  ctx->tl_outside = outside;
  ctx->tl_id_list = a->sorted_list; ctx->tl_id_nop = a->number_of_partners; ctx->tl_id_ptr = a->pointer;
  ctx->tl_pn = pn; ctx->tl_r0 = 0; ctx->tl_r1 = pn;
  ctx->graph_loop = -1;
  return LJ_OK;
}

extern "C" int lj_list_mirror(lj_ctx* ctx, const void* q, int64_t pn, int32_t layout, int64_t plane_stride,
                              double search_len, const int32_t* number_of_partners, const void* pointer,
                              int32_t pointer64, const int32_t* sorted_list, int64_t list_entries, int32_t flags,
                              int64_t* rows_outside_out, void* stream) {
  LJ_ENTER(ctx);
  LJ_REQUIRE(ctx, q && number_of_partners && pointer && sorted_list, "lj_list_mirror: null array");
  LJ_REQUIRE(ctx, pn > 0 && pn < 2147483647LL && search_len > 0.0 && list_entries >= 0, "lj_list_mirror: bad size");
  LJ_REQUIRE(ctx, layout == LJ_AOS_D3 || layout == LJ_AOS_D4 || layout == LJ_SOA_D, "lj_list_mirror: FP64 layouts only");
  if (layout == LJ_SOA_D) LJ_REQUIRE(ctx, plane_stride >= pn, "lj_list_mirror: SoA plane_stride < particle_number");
  if (layout == LJ_AOS_D4) LJ_REQUIRE(ctx, (uintptr_t)q % 32 == 0, "lj_list_mirror: double4 array must be 32-byte aligned");
  cudaStream_t st = lj_stream(ctx, stream);
  lj_list_args a{};
  a.q = q; a.pn = pn; a.layout = layout; a.plane_stride = plane_stride; a.search_len = search_len;
  a.number_of_partners = const_cast<int32_t*>(number_of_partners); a.pointer = const_cast<void*>(pointer);
  a.sorted_list = const_cast<int32_t*>(sorted_list); a.capacity = list_entries; a.pointer64 = pointer64; a.flags = flags;
  switch (layout) {
    case LJ_AOS_D3: return list_mirror_impl<LJ_AOS_D3>(ctx, &a, rows_outside_out, st);
    case LJ_AOS_D4: return list_mirror_impl<LJ_AOS_D4>(ctx, &a, rows_outside_out, st);
    default: return list_mirror_impl<LJ_SOA_D>(ctx, &a, rows_outside_out, st);
  }
}

extern "C" uint64_t lj_list_mirror_token(lj_ctx* ctx) { return (ctx && ctx->tl_valid) ? ctx->tl_token : 0; }

// positions of this step in cell order (the force kernel's TMA source).  part 0: every particle;
// 1 / 2 (lj_force_step_part): only the particles inside / outside the mirror's row range, and
// block 0 compacts the columns the force kernel that follows will walk (tile_select_columns).
int lj_celltile_permute(lj_ctx* ctx, const lj_force_args* a, cudaStream_t st, int part) {
  const unsigned blocks = (unsigned)blocks_for(a->pn, 256);
  tile_part tp{};
  tp.part = part; tp.r0 = ctx->tl_r0; tp.r1 = ctx->tl_r1;
  tp.cols = ctx->tl_cols; tp.ncols_all = ctx->tl_g.ntx * ctx->tl_g.nz;
  tp.zflag = ctx->tl_zflag; tp.cols_sel = ctx->tl_cols_sel;
  if (a->precision == LJ_PREC_MIXED) {
    if (ctx->tl_qfx_cap < a->pn) {
      if (ctx->tl_qfx) LJ_CUDA(ctx, cudaFreeAsync(ctx->tl_qfx, st));
      ctx->tl_qfx = nullptr;
      LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->tl_qfx, sizeof(int4) * (size_t)(a->pn + 2), ctx->pool, st));
      LJ_CUDA(ctx, cudaMemsetAsync(ctx->tl_qfx, 0, sizeof(int4) * (size_t)(a->pn + 2), st));
      ctx->tl_qfx_cap = a->pn;
      ctx->graph_loop = -1;  // a cached CUDA graph may hold the old pointer
    }
    const double scale = lj_fx_frame_for(a->cl2).scale;
    switch (a->layout) {
      case LJ_AOS_D3: k_tile_permute_fx<LJ_AOS_D3><<<blocks, 256, 0, st>>>(a->q, a->plane_stride, ctx->tl_order, a->pn, scale, ctx->tl_qfx, ctx->tl_geom, tp); break;
      case LJ_AOS_D4: k_tile_permute_fx<LJ_AOS_D4><<<blocks, 256, 0, st>>>(a->q, a->plane_stride, ctx->tl_order, a->pn, scale, ctx->tl_qfx, ctx->tl_geom, tp); break;
      default: k_tile_permute_fx<LJ_SOA_D><<<blocks, 256, 0, st>>>(a->q, a->plane_stride, ctx->tl_order, a->pn, scale, ctx->tl_qfx, ctx->tl_geom, tp); break;
    }
    LJ_LAUNCHED(ctx);
    return LJ_OK;
  }
  switch (a->layout) {
    case LJ_AOS_D3: k_tile_permute<LJ_AOS_D3><<<blocks, 256, 0, st>>>(a->q, a->plane_stride, ctx->tl_order, a->pn, reinterpret_cast<double2*>(ctx->tl_qs), ctx->tl_qz, ctx->tl_geom, tp); break;
    case LJ_AOS_D4: k_tile_permute<LJ_AOS_D4><<<blocks, 256, 0, st>>>(a->q, a->plane_stride, ctx->tl_order, a->pn, reinterpret_cast<double2*>(ctx->tl_qs), ctx->tl_qz, ctx->tl_geom, tp); break;
    default: k_tile_permute<LJ_SOA_D><<<blocks, 256, 0, st>>>(a->q, a->plane_stride, ctx->tl_order, a->pn, reinterpret_cast<double2*>(ctx->tl_qs), ctx->tl_qz, ctx->tl_geom, tp); break;
  }
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

int lj_sort_rows_launch(lj_ctx* ctx, int32_t* list, const int32_t* nop, const void* pointer,
                        int pointer64, int64_t pn, int64_t capacity, cudaStream_t st);

extern "C" int lj_build_list(lj_ctx* ctx, const lj_list_args* a, int64_t* number_of_pairs_out,
                             void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, a != nullptr, "lj_build_list: null args");
  LJ_REQUIRE(ctx, a->pn >= 0 && a->pn < 2147483647LL, "lj_build_list: particle_number out of range");
  LJ_REQUIRE(ctx, a->search_len > 0.0, "lj_build_list: search length must be positive");
  LJ_REQUIRE(ctx, a->capacity >= 0, "lj_build_list: negative capacity");
  cudaStream_t st = lj_stream(ctx, stream);
  if (a->pn == 0) {
    if (number_of_pairs_out) *number_of_pairs_out = 0;
    return LJ_OK;
  }
  LJ_REQUIRE(ctx, a->q && a->number_of_partners && a->pointer && (a->sorted_list || a->capacity == 0),
             "lj_build_list: null array");
  if (a->layout == LJ_SOA_D)
    LJ_REQUIRE(ctx, a->plane_stride >= a->pn, "lj_build_list: SoA plane_stride < particle_number");
  int rc;
  const bool tiles = (a->flags & LJ_LIST_TILES) && !a->half && a->layout != LJ_AOS_F4 && a->layout != LJ_AOS_F3;
  bool deferred = tiles;  // let the cell-tile fill pass write sorted_list too, if the engine allows
  // the tile engine (one search per build) serves the plain LJ_LIST_TILES build; when it does not apply it
  // leaves the binning done and the round-1 engine takes over
  bool binned = false;
  if (tiles && !(a->flags & (LJ_LIST_CLUSTERS | LJ_LIST_PER_PARTICLE_SEARCH)) && !lj_diag_set("LJ_TILE_OLD_ENGINE")) {
    bool done = false;
    switch (a->layout) {
      case LJ_AOS_D3: rc = build_list_tiles<LJ_AOS_D3>(ctx, a, st, &done); break;
      case LJ_AOS_D4:
        LJ_REQUIRE(ctx, (uintptr_t)a->q % 32 == 0, "lj_build_list: double4 array must be 32-byte aligned");
        rc = build_list_tiles<LJ_AOS_D4>(ctx, a, st, &done);
        break;
      default: rc = build_list_tiles<LJ_SOA_D>(ctx, a, st, &done); break;
    }
    if (rc) return rc;
    if (done) {
      if (a->flags & LJ_LIST_SORT_ROWS) {
        // sorting the public rows would leave the mirror in another order: still the same set per row,
        // and the mirror only ever is consumed by the cell-tile kernel
        rc = lj_sort_rows_launch(ctx, a->sorted_list, a->number_of_partners, a->pointer, a->pointer64,
                                 a->pn, a->capacity, st);
        if (rc) return rc;
      }
      if (number_of_pairs_out) return lj_list_result(ctx, number_of_pairs_out, nullptr, stream);
      return LJ_OK;
    }
    binned = true;
  }
  switch (a->layout) {
    case LJ_AOS_D3: rc = build_list_impl<LJ_AOS_D3>(ctx, a, st, &deferred, binned); break;
    case LJ_AOS_D4:
      LJ_REQUIRE(ctx, (uintptr_t)a->q % 32 == 0, "lj_build_list: double4 array must be 32-byte aligned");
      rc = build_list_impl<LJ_AOS_D4>(ctx, a, st, &deferred, binned);
      break;
    case LJ_SOA_D: rc = build_list_impl<LJ_SOA_D>(ctx, a, st, &deferred, binned); break;
    case LJ_AOS_F4:
      LJ_REQUIRE(ctx, (uintptr_t)a->q % 16 == 0, "lj_build_list: float4 array must be 16-byte aligned");
      rc = build_list_impl<LJ_AOS_F4>(ctx, a, st, &deferred);
      break;
    case LJ_AOS_F3:
      LJ_REQUIRE(ctx, (uintptr_t)a->q % 4 == 0, "lj_build_list: float3 array must be 4-byte aligned");
      rc = build_list_impl<LJ_AOS_F3>(ctx, a, st, &deferred);
      break;
    default: return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_build_list", "unknown layout");
  }
  if (rc) return rc;
  if (tiles) {
    bool filled = false;
    rc = build_tile_mirror(ctx, a, st, deferred, &filled);
    if (rc) return rc;
    if (deferred && !filled) {  // no mirror after all: the regular FILL pass
      switch (a->layout) {
        case LJ_AOS_D3: rc = cluster_fill<LJ_AOS_D3>(ctx, a, st, false); break;
        case LJ_AOS_D4: rc = cluster_fill<LJ_AOS_D4>(ctx, a, st, false); break;
        default: rc = cluster_fill<LJ_SOA_D>(ctx, a, st, false); break;
      }
      if (rc) return rc;
    }
  }
  if (a->flags & LJ_LIST_SORT_ROWS) {
    rc = lj_sort_rows_launch(ctx, a->sorted_list, a->number_of_partners, a->pointer, a->pointer64,
                             a->pn, a->capacity, st);
    if (rc) return rc;
  }
  if (number_of_pairs_out) return lj_list_result(ctx, number_of_pairs_out, nullptr, stream);
  return LJ_OK;
}

extern "C" int lj_list_result(lj_ctx* ctx, int64_t* number_of_pairs_out, int32_t* max_partners_out,
                              void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, ctx->totals != nullptr, "lj_list_result: no list has been built on this context");
  cudaStream_t st = lj_stream(ctx, stream);
  LJ_CUDA(ctx, cudaMemcpyAsync(ctx->totals_host, ctx->totals, sizeof(lj_list_totals),
                               cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  if (number_of_pairs_out) *number_of_pairs_out = (int64_t)ctx->totals_host->total;
  if (max_partners_out) *max_partners_out = ctx->totals_host->max_np;
  if (ctx->totals_host->overflow & 8)
    return lj_set_error(ctx, LJ_ERR_INVALID_LIST, "lj_build_list", "cell-tile mirror disagrees with the CSR list");
  if (ctx->totals_host->overflow & 2)
    return lj_set_error(ctx, LJ_ERR_OVERFLOW32, "lj_build_list", "list offsets exceed 32 bits: pass pointer64=1");
  if (ctx->totals_host->overflow & 1)
    return lj_set_error(ctx, LJ_ERR_CAPACITY, "lj_build_list", "sorted_list capacity too small (needed total returned)");
  return LJ_OK;
}

extern "C" int lj_build_ell(lj_ctx* ctx, const int32_t* sorted_list, const int32_t* nop,
                            const void* pointer, int32_t pointer64, int64_t pn, int32_t* tl,
                            int64_t capacity_entries, int32_t* max_partners_out, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, pn >= 0, "lj_build_ell: negative particle_number");
  cudaStream_t st = lj_stream(ctx, stream);
  if (pn == 0) { if (max_partners_out) *max_partners_out = 0; return LJ_OK; }
  LJ_REQUIRE(ctx, sorted_list && nop && pointer && tl, "lj_build_ell: null array");
  int rc = lj_scratch_reserve(ctx, 1, st);
  if (rc) return rc;
  int* mx = &ctx->totals->max_np;
  LJ_CUDA(ctx, cudaMemsetAsync(mx, 0, sizeof(int), st));
  k_max_np<<<4 * ctx->sm_count, 256, 0, st>>>(nop, pn, mx);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(&ctx->totals_host->max_np, mx, sizeof(int), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  const int max_np = ctx->totals_host->max_np;
  if (max_partners_out) *max_partners_out = max_np;
  if ((int64_t)max_np * pn > capacity_entries)
    return lj_set_error(ctx, LJ_ERR_CAPACITY, "lj_build_ell", "transposed_list capacity < max_partners*pn");
  if (max_np == 0) return LJ_OK;
  const unsigned blocks = (unsigned)blocks_for((int64_t)max_np * pn, 256);
  if (pointer64) k_csr_to_ell<true><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, max_np, tl);
  else k_csr_to_ell<false><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, max_np, tl);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

// CSR -> row-major padded table (one thread per table entry, zero padding)
template <bool PTR64>
__global__ void __launch_bounds__(256)
k_csr_to_ellrows(const int32_t* __restrict__ list, const int32_t* __restrict__ nop, const void* __restrict__ pointer,
                 int64_t pn, int width, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pn * width) return;
  const int64_t i = t / width;
  const int k = (int)(t - i * width);
  out[t] = k < nop[i] ? list[row_offset<PTR64>(pointer, i) + k] : 0;
}

extern "C" int lj_build_ell_rows(lj_ctx* ctx, const int32_t* sorted_list, const int32_t* nop,
                                 const void* pointer, int32_t pointer64, int64_t pn, int32_t width,
                                 int32_t* out, int64_t capacity_entries, int32_t* max_partners_out,
                                 void* stream) {
  LJ_ENTER(ctx);
  LJ_REQUIRE(ctx, pn >= 0 && width >= 0, "lj_build_ell_rows: negative size");
  cudaStream_t st = lj_stream(ctx, stream);
  if (pn == 0) { if (max_partners_out) *max_partners_out = 0; return LJ_OK; }
  LJ_REQUIRE(ctx, sorted_list && nop && pointer && out, "lj_build_ell_rows: null array");
  int rc = lj_scratch_reserve(ctx, 1, st);
  if (rc) return rc;
  int* mx = &ctx->totals->max_np;
  LJ_CUDA(ctx, cudaMemsetAsync(mx, 0, sizeof(int), st));
  k_max_np<<<4 * ctx->sm_count, 256, 0, st>>>(nop, pn, mx);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(&ctx->totals_host->max_np, mx, sizeof(int), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  const int max_np = ctx->totals_host->max_np;
  if (max_partners_out) *max_partners_out = max_np;
  // the reference's fixed NUM_NEIGH = 60 (cuda/force_cuda.cu:15) is smaller than its own longest row
  // (78 at rho = 0.5): its rows overlap.  Here a width that does not hold the longest row is an error.
  if (width < max_np)
    return lj_set_error(ctx, LJ_ERR_CAPACITY, "lj_build_ell_rows", "width < max_partners: rows would overlap");
  if ((int64_t)width * pn > capacity_entries)
    return lj_set_error(ctx, LJ_ERR_CAPACITY, "lj_build_ell_rows", "sorted_list2d capacity < width*pn");
  if (width == 0) return LJ_OK;
  const unsigned blocks = (unsigned)blocks_for((int64_t)width * pn, 256);
  if (pointer64) k_csr_to_ellrows<true><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, width, out);
  else k_csr_to_ellrows<false><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, width, out);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

extern "C" int lj_shuffle_rows(lj_ctx* ctx, int32_t* sorted_list, const int32_t* nop,
                               const void* pointer, int32_t pointer64, int64_t pn, uint32_t seed,
                               void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (pn <= 0) return LJ_OK;
  LJ_REQUIRE(ctx, sorted_list && nop && pointer, "lj_shuffle_rows: null array");
  // the cluster mirror stays correct as a SET, but drop it so that a shuffled list really is
  // consumed in its shuffled order
  if (ctx->cl_id_list == sorted_list) ctx->cl_valid = false;
  if (ctx->tl_id_list == sorted_list) ctx->tl_valid = false;
  ctx->graph_loop = -1;  // a cached CUDA graph may have captured the kernel that ran on the mirror
  cudaStream_t st = lj_stream(ctx, stream);
  const unsigned blocks = (unsigned)blocks_for(pn, 256);
  if (pointer64) k_shuffle_rows<true><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, seed);
  else k_shuffle_rows<false><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, seed);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

extern "C" int lj_validate_list(lj_ctx* ctx, const int32_t* sorted_list, const int32_t* nop,
                                const void* pointer, int32_t pointer64, int64_t pn,
                                int64_t number_of_pairs, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (pn <= 0) return LJ_OK;
  LJ_REQUIRE(ctx, sorted_list && nop && pointer, "lj_validate_list: null array");
  cudaStream_t st = lj_stream(ctx, stream);
  int rc = lj_scratch_reserve(ctx, 1, st);
  if (rc) return rc;
  int* bad = &ctx->totals->overflow;
  LJ_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(int), st));
  const int64_t n = pn > number_of_pairs ? pn : number_of_pairs;
  const unsigned blocks = (unsigned)blocks_for(n, 256);
  if (pointer64) k_validate<true><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, number_of_pairs, bad);
  else k_validate<false><<<blocks, 256, 0, st>>>(sorted_list, nop, pointer, pn, number_of_pairs, bad);
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(&ctx->totals_host->overflow, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  if (ctx->totals_host->overflow)
    return lj_set_error(ctx, LJ_ERR_INVALID_LIST, "lj_validate_list",
                        (ctx->totals_host->overflow & 1) ? "number_of_partners/pointer out of range"
                                                         : "sorted_list entry out of range");
  return LJ_OK;
}

// ------------------------------------------------------------------ optional row sort --
// LJ_LIST_SORT_ROWS: rows ascending in j, the order makepair() itself produces
// (cuda/force_cuda.cu:138-162).  One warp per row, rank sort out of shared memory (entries
// of a row are distinct, so ranks are a permutation).
namespace {
constexpr int kSortCap = 1024;  // entries per row held in shared memory

template <bool PTR64>
__global__ void __launch_bounds__(256)
k_sort_rows(int32_t* __restrict__ list, const int32_t* __restrict__ nop,
            const void* __restrict__ pointer, int64_t pn, int64_t capacity) {
  __shared__ int32_t buf[8][kSortCap];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + w;
  if (i >= pn) return;
  const int n = nop[i];
  const int64_t base = row_offset<PTR64>(pointer, i);
  if (n < 2 || base + n > capacity) return;
  int32_t* row = list + base;
  if (n <= kSortCap) {
    for (int k = lane; k < n; k += 32) buf[w][k] = row[k];
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const int32_t v = buf[w][k];
      int rank = 0;
      for (int m = 0; m < n; m++) rank += buf[w][m] < v;
      row[rank] = v;
    }
  } else if (lane == 0) {  // pathological row length: in-place insertion sort by one lane
    for (int k = 1; k < n; k++) {
      const int32_t v = row[k];
      int m = k - 1;
      while (m >= 0 && row[m] > v) { row[m + 1] = row[m]; m--; }
      row[m + 1] = v;
    }
  }
}
}  // namespace

int lj_sort_rows_launch(lj_ctx* ctx, int32_t* list, const int32_t* nop, const void* pointer,
                        int pointer64, int64_t pn, int64_t capacity, cudaStream_t st) {
  const unsigned blocks = (unsigned)((pn + 7) / 8);
  if (pointer64) k_sort_rows<true><<<blocks, 256, 0, st>>>(list, nop, pointer, pn, capacity);
  else k_sort_rows<false><<<blocks, 256, 0, st>>>(list, nop, pointer, pn, capacity);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}
