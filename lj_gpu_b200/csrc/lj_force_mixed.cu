// lj_force_mixed.cu -- FP32 pair arithmetic on 32-bit fixed-point coordinates, FP64 momenta.
//
// The reference allocates float3/float4 copies of q and p but never times them
// (cuda/force_cuda.cu:24-25,355-375).  Plain float positions cannot meet the 1e-5 parity bound
// against cpu_ref (a coordinate of ~50 sigma carries 2e-6 of rounding, amplified ~14x by the
// r^-14 force law), so the mixed mode keeps the caller's FP64 q/p interface and derives, every
// step, a 16-byte fixed-point copy: each coordinate is (x - centre)/unit rounded to int32 with
// unit = half-extent / 2^30.  A coordinate DIFFERENCE is then an exact integer subtraction; only
// its int->float conversion rounds (2^-24 relative).  All pair arithmetic is FP32 (MUFU.RCP
// reciprocal); per-lane FP32 partial sums are widened to FP64 before the shuffle reduction and
// the update of p.  The cutoff decision of pairs closer to r2 == CL2 than the representation
// error is re-taken in FP64 from the original positions, so no pair is ever misclassified
// (the unshifted LJ force jumps by 1.1e-2 at the cutoff).
//
// Three launches per step: bounding box, fixed-point conversion, force.
#include "lj_common.cuh"

namespace {

struct fix_params {
  double cx, cy, cz;  // centre of the bounding box
  double inv_unit;    // counts per length
  float unit;         // length per count
  float margin;       // bound on |r2_f32 - r2_f64| near the cutoff
};

__global__ void k_fix_setup(const unsigned long long* bb, double cutoff, fix_params* fp) {
  double lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = dec_ordered(bb[d]); hi[d] = dec_ordered(bb[3 + d]); }
  double E = fmax(hi[0] - lo[0], fmax(hi[1] - lo[1], hi[2] - lo[2]));
  if (!(E > 0.0)) E = 1.0;
  const double H = 0.5 * E * (1.0 + 1e-6);
  const double unit = H / 1073741824.0;  // 2^30 counts per half extent
  fp->cx = 0.5 * (lo[0] + hi[0]); fp->cy = 0.5 * (lo[1] + hi[1]); fp->cz = 0.5 * (lo[2] + hi[2]);
  fp->inv_unit = 1.0 / unit;
  fp->unit = (float)unit;
  // per-component error of a difference: one count (two half-count roundings) plus the
  // int->float rounding 2^-24 |d|, |d| <= cutoff*(1+eps) for pairs near the cutoff;
  // r2 error <= 3*(2|d| delta + delta^2) + four FP32 roundings of ~r2.  Doubled.
  const double u = 5.9604644775390625e-8;
  const double dmax = 1.01 * cutoff;
  const double delta = unit + u * dmax;
  const double m = 3.0 * (2.0 * dmax * delta + delta * delta) + 6.0 * u * dmax * dmax;
  fp->margin = (float)(2.0 * m);
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_to_fixed(const void* __restrict__ q, int64_t pn, int64_t plane, const fix_params* __restrict__ fp,
           int4* __restrict__ q32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double x, y, z;
  load_pos<LAYOUT>(q, i, plane, x, y, z);
  const double s = fp->inv_unit;
  q32[i] = make_int4(__double2int_rn((x - fp->cx) * s), __double2int_rn((y - fp->cy) * s),
                     __double2int_rn((z - fp->cz) * s), 0);
}

__device__ __forceinline__ float rcp_f32(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

constexpr int kUnroll = 4;

template <int G, int LAYOUT, bool PTR64>
__global__ void __launch_bounds__(1024)
lj_gather_mixed(const void* __restrict__ q, const int4* __restrict__ q32, void* __restrict__ p,
                int64_t row_begin, int64_t row_end, int64_t plane, float c24, float c48, float cl2f,
                double cl2, const fix_params* __restrict__ fp, const int32_t* __restrict__ list,
                const int32_t* __restrict__ nop, const void* __restrict__ pointer) {
  const int rows_per_block = blockDim.x / G;
  const int64_t i = row_begin + (int64_t)blockIdx.x * rows_per_block + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  if (i >= row_end) return;
  const float unit = fp->unit, margin = fp->margin;
  const float unit2 = unit * unit;
  const float lo_c = cl2f - margin, hi_c = cl2f + margin;
  const int4 me = __ldg(q32 + i);
  const int np = __ldg(nop + i);
  const int32_t* __restrict__ row = list + row_offset<PTR64>(pointer, i);
  float fx = 0.f, fy = 0.f, fz = 0.f;

  auto pair = [&](int j, int4 pj) {
    const float dx = (float)(pj.x - me.x), dy = (float)(pj.y - me.y), dz = (float)(pj.z - me.z);
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) * unit2;
    const float x = rcp_f32(r2);
    const float x3 = x * x * x;
    const float t = fmaf(-c48, x3, c24);
    const float df = (x * x3) * (t * unit);  // * unit: dx below is in counts
    bool in = r2 <= lo_c;
    if (!in && r2 < hi_c) {  // too close to the cutoff for FP32 to call: decide in FP64
      double xi, yi, zi, xj, yj, zj;
      load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
      load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
      const double ex = xj - xi, ey = yj - yi, ez = zj - zi;
      in = fma(ez, ez, fma(ey, ey, ex * ex)) <= cl2;
    }
    if (in) {
      fx = fmaf(df, dx, fx);
      fy = fmaf(df, dy, fy);
      fz = fmaf(df, dz, fz);
    }
  };

  int k = lg;
  for (; k + (kUnroll - 1) * G < np; k += kUnroll * G) {
    int j[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) j[u] = __ldg(row + k + u * G);
    int4 pj[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) pj[u] = __ldg(q32 + j[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; u++) pair(j[u], pj[u]);
  }
  for (; k < np; k += G) {
    const int j = __ldg(row + k);
    pair(j, __ldg(q32 + j));
  }
  double sx = (double)fx, sy = (double)fy, sz = (double)fz;
  if (G > 1) {
    sx = group_sum<G>(sx);
    sy = group_sum<G>(sy);
    sz = group_sum<G>(sz);
  }
  if (lg == 0) add_mom<LAYOUT>(p, i, plane, sx, sy, sz);
}

template <int G, int LAYOUT, bool PTR64>
void launch_mixed(const lj_force_args* a, const int4* q32, const fix_params* fp, int64_t r0, int64_t r1,
                  int tb, cudaStream_t st) {
  const int rows_per_block = tb / G;
  const unsigned blocks = (unsigned)((r1 - r0 + rows_per_block - 1) / rows_per_block);
  lj_gather_mixed<G, LAYOUT, PTR64><<<blocks, tb, 0, st>>>(
      a->q, q32, a->p, r0, r1, a->plane_stride, (float)(24.0 * a->dt), (float)(48.0 * a->dt),
      (float)a->cl2, a->cl2, fp, a->list, a->number_of_partners, a->pointer);
}

template <int LAYOUT, bool PTR64>
bool launch_mixed_g(int g, const lj_force_args* a, const int4* q32, const fix_params* fp, int64_t r0,
                    int64_t r1, int tb, cudaStream_t st) {
  switch (g) {
    case 1: launch_mixed<1, LAYOUT, PTR64>(a, q32, fp, r0, r1, tb, st); return true;
    case 2: launch_mixed<2, LAYOUT, PTR64>(a, q32, fp, r0, r1, tb, st); return true;
    case 4: launch_mixed<4, LAYOUT, PTR64>(a, q32, fp, r0, r1, tb, st); return true;
    case 8: launch_mixed<8, LAYOUT, PTR64>(a, q32, fp, r0, r1, tb, st); return true;
    case 16: launch_mixed<16, LAYOUT, PTR64>(a, q32, fp, r0, r1, tb, st); return true;
    case 32: launch_mixed<32, LAYOUT, PTR64>(a, q32, fp, r0, r1, tb, st); return true;
  }
  return false;
}

// --------------------------------------------------------------------------------------
// Mixed precision on the CLUSTER PAIR LIST: one warp per cluster of four rows, 32 consecutive
// union entries per trip, each lane gathers its fixed-point q[j] once (LDG.128) and evaluates it
// against the four members held in registers (12 ints + 12 FP32 accumulators: ~64 registers, so
// unlike the FP64 register-blocked kernel the occupancy stays high).  The FP32 pipe has four
// times the FP64 rate, so the 1.5x extra pair evaluations of the union are cheap here, while
// the L1/LSU cost per real pair drops 2.6x.
// --------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(128)
lj_gather_cluster_mixed(const void* __restrict__ q, const int4* __restrict__ q32, void* __restrict__ p,
                        int64_t row0, int64_t row_end, int64_t c_begin, int64_t c_end, int64_t plane,
                        float c24, float c48, float cl2f, double cl2, const fix_params* __restrict__ fp,
                        const uint32_t* __restrict__ cl_list, const long long* __restrict__ cl_ptr) {
  const int64_t c = c_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (c >= c_end) return;
  const int lane = threadIdx.x & 31;
  const long long off = __ldg(cl_ptr + c);
  const int U = (int)(__ldg(cl_ptr + c + 1) - off);
  const uint32_t* __restrict__ src = cl_list + off;
  const int64_t i0 = row0 + 4 * c;
  const unsigned self = (unsigned)i0;
  const float unit = fp->unit, margin = fp->margin;
  const float unit2 = unit * unit;
  const float lo_c = cl2f - margin, hi_c = cl2f + margin;
  int mx[4], my[4], mz[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int4 m = __ldg(q32 + (i0 + r < row_end ? i0 + r : i0));
    mx[r] = m.x; my[r] = m.y; mz[r] = m.z;
  }
  float ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0}, az[4] = {0, 0, 0, 0};

  auto eval = [&](uint32_t e, int4 pj) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const float dx = (float)(pj.x - mx[r]), dy = (float)(pj.y - my[r]), dz = (float)(pj.z - mz[r]);
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) * unit2;
      const float x = rcp_f32(r2);
      const float x3 = x * x * x;
      const float t = fmaf(-c48, x3, c24);
      const bool listed = (e >> (28 + r)) & 1u;
      bool in = r2 <= lo_c;
      if (listed && !in && r2 < hi_c) {  // too close to the cutoff for FP32 to call: decide in FP64
        double xi, yi, zi, xj, yj, zj;
        load_pos<LAYOUT>(q, i0 + r, plane, xi, yi, zi);
        load_pos<LAYOUT>(q, e & 0x0fffffffu, plane, xj, yj, zj);
        const double ex = xj - xi, ey = yj - yi, ez = zj - zi;
        in = fma(ez, ez, fma(ey, ey, ex * ex)) <= cl2;
      }
      const float df = (listed && in) ? (x * x3) * (t * unit) : 0.f;
      ax[r] = fmaf(df, dx, ax[r]);
      ay[r] = fmaf(df, dy, ay[r]);
      az[r] = fmaf(df, dz, az[r]);
    }
  };
  auto fetch = [&](int k, uint32_t& e, int4& pj) {
    e = k < U ? __ldg(src + k) : self;  // past the end: mask 0
    pj = __ldg(q32 + (e & 0x0fffffffu));
  };
  uint32_t ea, eb = 0;
  int4 pa, pb = make_int4(0, 0, 0, 0);
  fetch(lane, ea, pa);
  if (32 < U) fetch(32 + lane, eb, pb);
  for (int k0 = 0; k0 < U; k0 += 64) {
    eval(ea, pa);
    if (k0 + 64 < U) fetch(k0 + 64 + lane, ea, pa);
    if (k0 + 32 < U) {
      eval(eb, pb);
      if (k0 + 96 < U) fetch(k0 + 96 + lane, eb, pb);
    }
  }
  // FP32 per-lane partial sums -> FP64 butterfly over the warp for the four members
#pragma unroll
  for (int r = 0; r < 4; r++) {
    double sx = (double)ax[r], sy = (double)ay[r], sz = (double)az[r];
    sx = group_sum<32>(sx); sy = group_sum<32>(sy); sz = group_sum<32>(sz);
    if (lane == r && i0 + r < row_end) add_mom<LAYOUT>(p, i0 + r, plane, sx, sy, sz);
  }
}

}  // namespace

bool lj_cluster_usable(const lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1);

int lj_force_mixed_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, int g, int tb,
                          cudaStream_t st) {
  const int64_t pn = a->pn;
  if (!ctx->bbox) {
    int rc = lj_scratch_reserve(ctx, 1, st);
    if (rc) return rc;
  }
  if (ctx->q32_len < pn) {
    if (ctx->q32) LJ_CUDA(ctx, cudaFreeAsync(ctx->q32, st));
    ctx->q32 = nullptr;
    LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->q32, sizeof(int4) * (size_t)pn + 256, ctx->pool, st));
    ctx->q32_len = pn;
  }
  // fix_params live in the second half of the 256-byte grid block (first half: cell grid)
  fix_params* fp = reinterpret_cast<fix_params*>(reinterpret_cast<char*>(ctx->grid) + 128);
  int4* q32 = reinterpret_cast<int4*>(ctx->q32);
  int rc = lj_bbox_launch(ctx, a->q, a->layout, pn, a->plane_stride, nullptr, st);
  if (rc) return rc;
  k_fix_setup<<<1, 1, 0, st>>>(reinterpret_cast<const unsigned long long*>(ctx->bbox), sqrt(a->cl2), fp);
  LJ_LAUNCHED(ctx);
  const unsigned cb = (unsigned)((pn + 255) / 256);
  switch (a->layout) {
    case LJ_AOS_D3: k_to_fixed<LJ_AOS_D3><<<cb, 256, 0, st>>>(a->q, pn, a->plane_stride, fp, q32); break;
    case LJ_AOS_D4: k_to_fixed<LJ_AOS_D4><<<cb, 256, 0, st>>>(a->q, pn, a->plane_stride, fp, q32); break;
    case LJ_AOS_F4: k_to_fixed<LJ_AOS_F4><<<cb, 256, 0, st>>>(a->q, pn, a->plane_stride, fp, q32); break;
    case LJ_AOS_F3: k_to_fixed<LJ_AOS_F3><<<cb, 256, 0, st>>>(a->q, pn, a->plane_stride, fp, q32); break;
    default: k_to_fixed<LJ_SOA_D><<<cb, 256, 0, st>>>(a->q, pn, a->plane_stride, fp, q32); break;
  }
  LJ_LAUNCHED(ctx);
  if (a->variant == LJ_VARIANT_CLUSTER) {
    const int64_t c0 = (r0 - ctx->cl_r0) / 4, c1 = (r1 - ctx->cl_r0 + 3) / 4;
    const unsigned blocks = (unsigned)((c1 - c0 + 3) / 4);
    const float c24 = (float)(24.0 * a->dt), c48 = (float)(48.0 * a->dt), cl2f = (float)a->cl2;
    switch (a->layout) {
      case LJ_AOS_D3:
        lj_gather_cluster_mixed<LJ_AOS_D3><<<blocks, 128, 0, st>>>(a->q, q32, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2f, a->cl2, fp, ctx->cl_list, ctx->cl_ptr);
        break;
      case LJ_AOS_D4:
        lj_gather_cluster_mixed<LJ_AOS_D4><<<blocks, 128, 0, st>>>(a->q, q32, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2f, a->cl2, fp, ctx->cl_list, ctx->cl_ptr);
        break;
      case LJ_AOS_F4:
        lj_gather_cluster_mixed<LJ_AOS_F4><<<blocks, 128, 0, st>>>(a->q, q32, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2f, a->cl2, fp, ctx->cl_list, ctx->cl_ptr);
        break;
      case LJ_AOS_F3:
        lj_gather_cluster_mixed<LJ_AOS_F3><<<blocks, 128, 0, st>>>(a->q, q32, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2f, a->cl2, fp, ctx->cl_list, ctx->cl_ptr);
        break;
      default:
        lj_gather_cluster_mixed<LJ_SOA_D><<<blocks, 128, 0, st>>>(a->q, q32, a->p, ctx->cl_r0, ctx->cl_r1, c0, c1, a->plane_stride, c24, c48, cl2f, a->cl2, fp, ctx->cl_list, ctx->cl_ptr);
        break;
    }
    LJ_LAUNCHED(ctx);
    return LJ_OK;
  }
  bool ok = false;
  switch (a->layout) {
    case LJ_AOS_D3:
      ok = a->pointer64 ? launch_mixed_g<LJ_AOS_D3, true>(g, a, q32, fp, r0, r1, tb, st)
                        : launch_mixed_g<LJ_AOS_D3, false>(g, a, q32, fp, r0, r1, tb, st);
      break;
    case LJ_AOS_D4:
      ok = a->pointer64 ? launch_mixed_g<LJ_AOS_D4, true>(g, a, q32, fp, r0, r1, tb, st)
                        : launch_mixed_g<LJ_AOS_D4, false>(g, a, q32, fp, r0, r1, tb, st);
      break;
    case LJ_SOA_D:
      ok = a->pointer64 ? launch_mixed_g<LJ_SOA_D, true>(g, a, q32, fp, r0, r1, tb, st)
                        : launch_mixed_g<LJ_SOA_D, false>(g, a, q32, fp, r0, r1, tb, st);
      break;
    case LJ_AOS_F4:
      ok = a->pointer64 ? launch_mixed_g<LJ_AOS_F4, true>(g, a, q32, fp, r0, r1, tb, st)
                        : launch_mixed_g<LJ_AOS_F4, false>(g, a, q32, fp, r0, r1, tb, st);
      break;
    case LJ_AOS_F3:
      ok = a->pointer64 ? launch_mixed_g<LJ_AOS_F3, true>(g, a, q32, fp, r0, r1, tb, st)
                        : launch_mixed_g<LJ_AOS_F3, false>(g, a, q32, fp, r0, r1, tb, st);
      break;
  }
  LJ_REQUIRE(ctx, ok, "lj_force_step: no mixed kernel for this configuration");
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}
