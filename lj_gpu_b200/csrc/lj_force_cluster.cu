// lj_force_cluster.cu -- FP64 force kernel on the cluster pair list (sm_100a).
//
// Why: ncu shows the per-row gather kernels bound by L1 data-pipe wavefronts (91 % of peak):
// every pair fetches its own 32-byte q[j] through L1, ~0.6 wavefronts per pair.  Four
// consecutive particles (one FCC cell in lattice order) share most of their neighbours, so the
// list build also emits, per cluster of four rows, the UNION of the rows with a 4-bit member
// mask per entry.  Here a warp takes 32 consecutive union entries, each lane gathers its q[j]
// ONCE and evaluates it against the four cluster members held in registers: 0.25 wavefronts
// and 0.38 list words per real pair; masked-out (member, j) combinations are neutralised with an
// impossible cutoff.  The price is FP64 work on combinations that are not listed pairs
// (x1.5 at rho = 1.0), i.e. the kernel trades L1 traffic for FP64 issue, which had 68 % headroom.
//
// Structure = lj_force_tile.cu: persistent CTAs, tiles of kClTile clusters whose entries are one
// contiguous segment staged by a TMA bulk copy (producer warp, mbarrier double buffering), eight
// consumer warps, four clusters each per tile, transposing butterfly for the 4 member sums.
// Results equal the per-row kernels up to summation order (rows are summed in union order).
#include "lj_common.cuh"
#include "lj_tile.cuh"

namespace {

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

constexpr int kClWarps = 8;
constexpr int kClThreads = kClWarps * 32 + 32;  // + producer warp
constexpr int kClPerWarp = 2;
constexpr int kClTile = kClWarps * kClPerWarp;  // clusters per tile (128 rows)
constexpr int kClCapInts = kClTile * 256;       // staged entries per tile (rho=1: ~200/cluster)

template <int LAYOUT>
__global__ void __launch_bounds__(kClThreads, 2)
lj_gather_cluster(const void* __restrict__ q, void* __restrict__ p, int64_t row0, int64_t row_end,
                  int64_t c_begin, int64_t c_end, int64_t plane, double c24, double c48,
                  long long cl2_bits, const uint32_t* __restrict__ cl_list,
                  const long long* __restrict__ cl_ptr, int64_t cl_entries) {
  extern __shared__ __align__(16) uint32_t stage_u[];  // 2 x kClCapInts
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
  __shared__ long long seg_base[2];
  __shared__ int seg_len[2];
  __shared__ double qi_s[kClWarps][kClPerWarp * 4][3];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; b++) {
      mbar_init(&full_bar[b], 1);
      mbar_init(&empty_bar[b], kClWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = (c_end - c_begin + kClTile - 1) / kClTile;

  if (warp == kClWarps) {  // ------------------------------------------------ producer warp
    if (lane == 0) {
      int n = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, n++) {
        const int b = n & 1;
        if (n >= 2) mbar_wait(&empty_bar[b], ((n >> 1) - 1) & 1);
        const int64_t cf = c_begin + t * kClTile;
        const int64_t cl = cf + kClTile < c_end ? cf + kClTile : c_end;
        const long long s0 = cl_ptr[cf], s1 = cl_ptr[cl];
        const long long s0a = s0 & ~3ll;  // 16-byte aligned start
        long long len = s1 - s0a;
        if (len > kClCapInts) len = kClCapInts;
        const long long up = (len + 3) & ~3ll;
        len = (up <= kClCapInts && s0a + up <= cl_entries) ? up : (len & ~3ll);
        seg_base[b] = s0a;
        seg_len[b] = (int)len;
        if (len > 0) {
          mbar_arrive_expect_tx(&full_bar[b], (uint32_t)(len * 4));
          bulk_g2s(stage_u + (size_t)b * kClCapInts, cl_list + s0a, (uint32_t)(len * 4), &full_bar[b]);
        } else {
          mbar_arrive(&full_bar[b]);
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------ consumer warps
  int n = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, n++) {
    const int b = n & 1;
    const int64_t cw = c_begin + t * kClTile + (int64_t)warp * kClPerWarp;  // my first cluster
    // my 16 member positions -> shared memory, my 5 cluster offsets -> registers; both are
    // in flight while the TMA copy of the tile's entries lands
    if (lane < kClPerWarp * 4) {
      const int64_t i = row0 + 4 * cw + lane;
      double x = 0.0, y = 0.0, z = 0.0;
      if (cw + lane / 4 < c_end && i < row_end) load_pos<LAYOUT>(q, i, plane, x, y, z);
      qi_s[warp][lane][0] = x; qi_s[warp][lane][1] = y; qi_s[warp][lane][2] = z;
    }
    long long my_ptr = 0;
    if (lane <= kClPerWarp) {
      const int64_t c = cw + lane < c_end ? cw + lane : c_end;
      my_ptr = cl_ptr[c];
    }
    __syncwarp();
    mbar_wait(&full_bar[b], (n >> 1) & 1);
    const long long sbase = seg_base[b];
    const int slen = seg_len[b];
    const uint32_t* __restrict__ sbuf = stage_u + (size_t)b * kClCapInts;

#pragma unroll 1
    for (int cc = 0; cc < kClPerWarp; cc++) {
      const int64_t c = cw + cc;
      const long long off = __shfl_sync(0xffffffffu, my_ptr, cc);
      const int U = (int)(__shfl_sync(0xffffffffu, my_ptr, cc + 1) - off);
      if (c >= c_end || U <= 0) continue;  // warp-uniform
      double xi[4], yi[4], zi[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        xi[r] = qi_s[warp][cc * 4 + r][0];
        yi[r] = qi_s[warp][cc * 4 + r][1];
        zi[r] = qi_s[warp][cc * 4 + r][2];
      }
      const long long rel = off - sbase;
      const bool in_smem = rel >= 0 && rel + U <= (long long)slen;
      const uint32_t* __restrict__ src = in_smem ? sbuf + rel : cl_list + off;
      const unsigned self = (unsigned)(row0 + 4 * c);  // a valid particle for idle lanes, mask 0
      double ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0}, az[4] = {0, 0, 0, 0};

      // software pipeline over 32-entry chunks: the next chunk's entry and q[j] are in flight
      // while the current chunk's four pair evaluations run
      auto fetch = [&](int k, uint32_t& e, double& x, double& y, double& z) {
        e = k < U ? (in_smem ? src[k] : __ldg(src + k)) : self;  // past the end: mask 0
        load_pos<LAYOUT>(q, e & 0x0fffffffu, plane, x, y, z);
      };
      // two chunks ahead, ping-pong register sets (no register moves on the critical path)
      auto eval = [&](uint32_t e, double xj, double yj, double zj) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const long long lim = ((e >> (28 + r)) & 1u) ? cl2_bits : -1ll;
          lj_pair(xj - xi[r], yj - yi[r], zj - zi[r], c24, c48, lim, ax[r], ay[r], az[r]);
        }
      };
      uint32_t ea, eb = 0;
      double xa, ya, za, xb = 0.0, yb = 0.0, zb = 0.0;
      fetch(lane, ea, xa, ya, za);
      if (32 < U) fetch(32 + lane, eb, xb, yb, zb);
      for (int k0 = 0; k0 < U; k0 += 64) {
        eval(ea, xa, ya, za);
        if (k0 + 64 < U) fetch(k0 + 64 + lane, ea, xa, ya, za);
        if (k0 + 32 < U) {
          eval(eb, xb, yb, zb);
          if (k0 + 96 < U) fetch(k0 + 96 + lane, eb, xb, yb, zb);
        }
      }
      int my_row;
      const double sx = batch_sum<32, 4>(ax, lane, 0xffffffffu, my_row);
      const double sy = batch_sum<32, 4>(ay, lane, 0xffffffffu, my_row);
      const double sz = batch_sum<32, 4>(az, lane, 0xffffffffu, my_row);
      const int64_t wrow = row0 + 4 * c + my_row;
      // one writer per member row; RED (no return value) so that the warp does not stall on a
      // load of p at the end of every cluster.  Exactly one add per component and step:
      // the result is deterministic.
      if ((lane & 7) == 0 && wrow < row_end) red_mom<LAYOUT>(p, wrow, plane, sx, sy, sz);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[b]);  // this warp no longer reads stage[b] / qi_s
  }
}

// --------------------------------------------------------------------------------------
// Lane-per-member variant: a warp = 8 entry slots x 4 cluster members.  The four lanes of a slot
// load the SAME q[j] (one sector request, coalesced by the LSU), every lane evaluates ONE pair per
// entry like the per-row kernel, so registers stay at the per-row level (~50, 58 % occupancy)
// instead of the 96 of the register-blocked kernel above, whose two resident CTAs cannot hide
// the gather latency (ncu: 27 % warps active, FP64 pipe 39 %).  No shared memory, no TMA: eight
// consecutive entries are one 32-byte sector.
// --------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(1024)
lj_gather_cluster_lanes(const void* __restrict__ q, void* __restrict__ p, int64_t row0, int64_t row_end,
                        int64_t c_begin, int64_t c_end, int64_t plane, double c24, double c48,
                        long long cl2_bits, const uint32_t* __restrict__ cl_list,
                        const long long* __restrict__ cl_ptr) {
  const int64_t c = c_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (c >= c_end) return;  // whole warps leave
  const int lane = threadIdx.x & 31;
  const int r = lane & 3, slot = lane >> 2;
  const int64_t i = row0 + 4 * c + r;
  const bool member = i < row_end;
  const int64_t iq = member ? i : row0 + 4 * c;  // absent member: compute on a valid particle, masked
  double xi, yi, zi;
  load_pos<LAYOUT>(q, iq, plane, xi, yi, zi);
  const long long off = __ldg(cl_ptr + c);
  const int U = (int)(__ldg(cl_ptr + c + 1) - off);
  const uint32_t* __restrict__ src = cl_list + off;
  const unsigned self = (unsigned)iq;
  const unsigned rbit = member ? (1u << (28 + r)) : 0u;

  double fx = 0.0, fy = 0.0, fz = 0.0;
  int k = slot;
  for (; k + 24 < U; k += 32) {  // four entries per lane, all valid
    uint32_t e[4];
#pragma unroll
    for (int u = 0; u < 4; u++) e[u] = __ldg(src + k + 8 * u);
    double xj[4], yj[4], zj[4];
#pragma unroll
    for (int u = 0; u < 4; u++) load_pos<LAYOUT>(q, e[u] & 0x0fffffffu, plane, xj[u], yj[u], zj[u]);
#pragma unroll
    for (int u = 0; u < 4; u++)
      lj_pair(xj[u] - xi, yj[u] - yi, zj[u] - zi, c24, c48, (e[u] & rbit) ? cl2_bits : -1ll, fx, fy, fz);
  }
  for (; k - slot < U; k += 8) {  // warp-uniform trip count; lanes past the end are masked
    const uint32_t e = k < U ? __ldg(src + k) : self;
    double xj, yj, zj;
    load_pos<LAYOUT>(q, e & 0x0fffffffu, plane, xj, yj, zj);
    lj_pair(xj - xi, yj - yi, zj - zi, c24, c48, (e & rbit) ? cl2_bits : -1ll, fx, fy, fz);
  }
#pragma unroll
  for (int m = 4; m <= 16; m <<= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, m);
    fy += __shfl_xor_sync(0xffffffffu, fy, m);
    fz += __shfl_xor_sync(0xffffffffu, fz, m);
  }
  if (slot == 0 && member) add_mom<LAYOUT>(p, i, plane, fx, fy, fz);
}

// --------------------------------------------------------------------------------------
// Register-blocked variant WITHOUT the tile machinery: one warp per cluster, entries read
// straight from global memory (32 consecutive words = one line), two chunks in flight.
// --------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(128)
lj_gather_cluster_rb(const void* __restrict__ q, void* __restrict__ p, int64_t row0, int64_t row_end,
                     int64_t c_begin, int64_t c_end, int64_t plane, double c24, double c48,
                     long long cl2_bits, const uint32_t* __restrict__ cl_list,
                     const long long* __restrict__ cl_ptr) {
  const int64_t c = c_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (c >= c_end) return;
  const int lane = threadIdx.x & 31;
  const long long off = __ldg(cl_ptr + c);
  const int U = (int)(__ldg(cl_ptr + c + 1) - off);
  const uint32_t* __restrict__ src = cl_list + off;
  const unsigned self = (unsigned)(row0 + 4 * c);
  auto fetch = [&](int k, uint32_t& e, double& x, double& y, double& z) {
    e = k < U ? __ldg(src + k) : self;
    load_pos<LAYOUT>(q, e & 0x0fffffffu, plane, x, y, z);
  };
  uint32_t ea, eb = 0;
  double xa, ya, za, xb = 0.0, yb = 0.0, zb = 0.0;
  fetch(lane, ea, xa, ya, za);
  if (32 < U) fetch(32 + lane, eb, xb, yb, zb);
  double xi[4], yi[4], zi[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int64_t i = row0 + 4 * c + r;
    load_pos<LAYOUT>(q, i < row_end ? i : row0 + 4 * c, plane, xi[r], yi[r], zi[r]);
  }
  double ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0}, az[4] = {0, 0, 0, 0};
  auto eval = [&](uint32_t e, double xj, double yj, double zj) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const long long lim = ((e >> (28 + r)) & 1u) ? cl2_bits : -1ll;
      lj_pair(xj - xi[r], yj - yi[r], zj - zi[r], c24, c48, lim, ax[r], ay[r], az[r]);
    }
  };
  for (int k0 = 0; k0 < U; k0 += 64) {
    eval(ea, xa, ya, za);
    if (k0 + 64 < U) fetch(k0 + 64 + lane, ea, xa, ya, za);
    if (k0 + 32 < U) {
      eval(eb, xb, yb, zb);
      if (k0 + 96 < U) fetch(k0 + 96 + lane, eb, xb, yb, zb);
    }
  }
  int my_row;
  const double sx = batch_sum<32, 4>(ax, lane, 0xffffffffu, my_row);
  const double sy = batch_sum<32, 4>(ay, lane, 0xffffffffu, my_row);
  const double sz = batch_sum<32, 4>(az, lane, 0xffffffffu, my_row);
  const int64_t wrow = row0 + 4 * c + my_row;
  if ((lane & 7) == 0 && wrow < row_end) red_mom<LAYOUT>(p, wrow, plane, sx, sy, sz);
}

template <int LAYOUT>
int launch_cluster_lanes(lj_ctx* ctx, const lj_force_args* a, int64_t c0, int64_t c1, double c24,
                         double c48, long long cl2_bits, int tb, cudaStream_t st) {
  const int64_t warps_per_block = tb / 32;
  const unsigned blocks = (unsigned)((c1 - c0 + warps_per_block - 1) / warps_per_block);
  lj_gather_cluster_lanes<LAYOUT><<<blocks, tb, 0, st>>>(a->q, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1,
                                                          a->plane_stride, c24, c48, cl2_bits,
                                                          ctx->cl_list, ctx->cl_ptr);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

template <int LAYOUT>
int launch_cluster(lj_ctx* ctx, const lj_force_args* a, int64_t c0, int64_t c1, double c24, double c48,
                   long long cl2_bits, cudaStream_t st) {
  const size_t smem = (size_t)2 * kClCapInts * sizeof(uint32_t);
  auto kern = lj_gather_cluster<LAYOUT>;
  // shared-memory opt-in and occupancy are per DEVICE: cached in the context, not in the process
  LJ_FUNC_SMEM(ctx, kern, smem);
  int& per_sm = ctx->func_occ[reinterpret_cast<const void*>(kern)];
  if (per_sm < 1) {
    LJ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kClThreads, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const int64_t ntiles = (c1 - c0 + kClTile - 1) / kClTile;
  int64_t grid = (int64_t)ctx->sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kClThreads, smem, st>>>(a->q, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1,
                                                  a->plane_stride, c24, c48, cl2_bits, ctx->cl_list,
                                                  ctx->cl_ptr, ctx->cl_entries);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

}  // namespace

// true when the cluster mirror describes exactly the list arrays of this call and the requested
// row range falls on cluster boundaries
bool lj_cluster_usable(const lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1) {
  if (!ctx->cl_valid || a->list_layout != LJ_LIST_CSR) return false;
  if (a->list != ctx->cl_id_list || a->number_of_partners != ctx->cl_id_nop ||
      a->pointer != ctx->cl_id_ptr || a->pn != ctx->cl_pn)
    return false;
  if (r0 < ctx->cl_r0 || r1 > ctx->cl_r1) return false;
  if ((r0 - ctx->cl_r0) % 4 != 0) return false;
  if (r1 != ctx->cl_r1 && (r1 - ctx->cl_r0) % 4 != 0) return false;
  return true;
}

int lj_force_cluster_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, double c24,
                            double c48, long long cl2_bits, cudaStream_t st) {
  const int64_t c0 = (r0 - ctx->cl_r0) / 4, c1 = (r1 - ctx->cl_r0 + 3) / 4;
  if (a->group == 16) {  // register-blocked, one warp per cluster, no tile staging
    const unsigned blocks = (unsigned)((c1 - c0 + 3) / 4);
    switch (a->layout) {
      case LJ_AOS_D4: lj_gather_cluster_rb<LJ_AOS_D4><<<blocks, 128, 0, st>>>(a->q, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2_bits, ctx->cl_list, ctx->cl_ptr); break;
      case LJ_AOS_D3: lj_gather_cluster_rb<LJ_AOS_D3><<<blocks, 128, 0, st>>>(a->q, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2_bits, ctx->cl_list, ctx->cl_ptr); break;
      default: lj_gather_cluster_rb<LJ_SOA_D><<<blocks, 128, 0, st>>>(a->q, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2_bits, ctx->cl_list, ctx->cl_ptr); break;
    }
    LJ_LAUNCHED(ctx);
    return LJ_OK;
  }
  if (a->group != 32) {  // default: lane-per-member; group = 32 selects the tiled register-blocked kernel
    const int tb = a->threads_per_block ? a->threads_per_block : 128;
    switch (a->layout) {
      case LJ_AOS_D4: return launch_cluster_lanes<LJ_AOS_D4>(ctx, a, c0, c1, c24, c48, cl2_bits, tb, st);
      case LJ_AOS_D3: return launch_cluster_lanes<LJ_AOS_D3>(ctx, a, c0, c1, c24, c48, cl2_bits, tb, st);
      case LJ_SOA_D: return launch_cluster_lanes<LJ_SOA_D>(ctx, a, c0, c1, c24, c48, cl2_bits, tb, st);
    }
  }
  switch (a->layout) {
    case LJ_AOS_D4: return launch_cluster<LJ_AOS_D4>(ctx, a, c0, c1, c24, c48, cl2_bits, st);
    case LJ_AOS_D3: return launch_cluster<LJ_AOS_D3>(ctx, a, c0, c1, c24, c48, cl2_bits, st);
    case LJ_SOA_D: return launch_cluster<LJ_SOA_D>(ctx, a, c0, c1, c24, c48, cl2_bits, st);
  }
  return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_force_step", "layout");
}
