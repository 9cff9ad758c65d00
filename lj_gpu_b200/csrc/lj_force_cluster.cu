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

constexpr int kClWarps = 8;
constexpr int kClThreads = kClWarps * 32 + 32;  // + producer warp
constexpr int kClPerWarp = 4;
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

      for (int k = lane; k < U; k += 64) {
        // two entries per lane and trip; the second may be past the end (mask 0 -> no effect)
        uint32_t e0, e1;
        const bool v1 = k + 32 < U;
        if (in_smem) { e0 = src[k]; e1 = v1 ? src[k + 32] : self; }
        else { e0 = __ldg(src + k); e1 = v1 ? __ldg(src + k + 32) : self; }
        double x0, y0, z0, x1, y1, z1;
        load_pos<LAYOUT>(q, e0 & 0x0fffffffu, plane, x0, y0, z0);
        load_pos<LAYOUT>(q, e1 & 0x0fffffffu, plane, x1, y1, z1);
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const long long lim = ((e0 >> (28 + r)) & 1u) ? cl2_bits : -1ll;
          lj_pair(x0 - xi[r], y0 - yi[r], z0 - zi[r], c24, c48, lim, ax[r], ay[r], az[r]);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const long long lim = ((e1 >> (28 + r)) & 1u) ? cl2_bits : -1ll;
          lj_pair(x1 - xi[r], y1 - yi[r], z1 - zi[r], c24, c48, lim, ax[r], ay[r], az[r]);
        }
      }
      int my_row;
      const double sx = batch_sum<32, 4>(ax, lane, 0xffffffffu, my_row);
      const double sy = batch_sum<32, 4>(ay, lane, 0xffffffffu, my_row);
      const double sz = batch_sum<32, 4>(az, lane, 0xffffffffu, my_row);
      const int64_t wrow = row0 + 4 * c + my_row;
      if ((lane & 7) == 0 && wrow < row_end) add_mom<LAYOUT>(p, wrow, plane, sx, sy, sz);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[b]);  // this warp no longer reads stage[b] / qi_s
  }
}

template <int LAYOUT>
int launch_cluster(lj_ctx* ctx, const lj_force_args* a, int64_t c0, int64_t c1, double c24, double c48,
                   long long cl2_bits, cudaStream_t st) {
  const size_t smem = (size_t)2 * kClCapInts * sizeof(uint32_t);
  auto kern = lj_gather_cluster<LAYOUT>;
  static bool configured = false;
  static int per_sm = 1;
  if (!configured) {
    LJ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LJ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kClThreads, smem));
    if (per_sm < 1) per_sm = 1;
    configured = true;
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
  if (!ctx->cl_valid || a->list_layout != LJ_LIST_CSR || a->precision != LJ_PREC_FP64) return false;
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
  switch (a->layout) {
    case LJ_AOS_D4: return launch_cluster<LJ_AOS_D4>(ctx, a, c0, c1, c24, c48, cl2_bits, st);
    case LJ_AOS_D3: return launch_cluster<LJ_AOS_D3>(ctx, a, c0, c1, c24, c48, cl2_bits, st);
    case LJ_SOA_D: return launch_cluster<LJ_SOA_D>(ctx, a, c0, c1, c24, c48, cl2_bits, st);
  }
  return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_force_step", "layout");
}
