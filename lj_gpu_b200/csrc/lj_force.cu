// lj_force.cu -- Lennard-Jones pair-force / momentum-update kernels for sm_100a.
//
// Replaces cuda/kernel.cuh (20 Kepler/Pascal kernels) behind lj_force_step().  One templated
// gather kernel covers the reference's thread-per-i ... warp-per-i spectrum through the lanes-
// per-row parameter G; a column-major ELL kernel covers the transposed_list variants; a
// half-list kernel with FP64 RED atomics covers the *_with_aar family; a CTA-tile kernel
// stages the j-indices of a block of rows in shared memory with one TMA bulk copy.
//
// Design notes (details and roofline in DESIGN.md):
//  * q[j] gather = one 32 B sector per pair for double4 (LDG.E.256), served by L1/L2 -- the
//    compulsory HBM stream is the int32 list (4 B/pair).
//  * pair body: 17 FP64-pipe instructions (see lj_pair), reciprocal by MUFU.RCP64H + 3 DFMA,
//    cutoff compare in the integer pipe, accumulation predicated.
//  * reductions by __shfl_xor over G lanes, no atomics in the gather kernels -> results are
//    bit-reproducible run to run.
#include "lj_common.cuh"

namespace {

constexpr int kUnroll = 4;

// --------------------------------------------------------------------------------------
// Gather on a CSR list, G lanes per row.  G=32 is the reference's warp_unroll mapping
// (cuda/kernel.cuh:821-904), G=1 its thread-per-i mapping (cuda/kernel.cuh:67-100).
// --------------------------------------------------------------------------------------
template <int G, int LAYOUT, bool PTR64>
__global__ void __launch_bounds__(1024)
lj_gather_csr(const void* __restrict__ q, void* __restrict__ p, int64_t row_begin, int64_t row_end,
              int64_t plane, double c24, double c48, long long cl2_bits,
              const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
              const void* __restrict__ pointer) {
  const int rows_per_block = blockDim.x / G;
  const int64_t i = row_begin + (int64_t)blockIdx.x * rows_per_block + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  // whole groups leave together (G divides 32), so the shuffles below stay convergent
  if (i >= row_end) return;

  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int32_t* __restrict__ row = list + row_offset<PTR64>(pointer, i);

  double fx = 0.0, fy = 0.0, fz = 0.0;
  int k = lg;
  // full tiles: kUnroll independent index loads, then kUnroll independent gathers
  for (; k + (kUnroll - 1) * G < np; k += kUnroll * G) {
    int j[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) j[u] = __ldg(row + k + u * G);
    double xj[kUnroll], yj[kUnroll], zj[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) load_pos<LAYOUT>(q, j[u], plane, xj[u], yj[u], zj[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; u++)
      lj_pair(xj[u] - xi, yj[u] - yi, zj[u] - zi, c24, c48, cl2_bits, fx, fy, fz);
  }
  for (; k < np; k += G) {
    const int j = __ldg(row + k);
    double xj, yj, zj;
    load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
    lj_pair(xj - xi, yj - yi, zj - zi, c24, c48, cl2_bits, fx, fy, fz);
  }

  if (G > 1) {
    fx = group_sum<G>(fx);
    fy = group_sum<G>(fy);
    fz = group_sum<G>(fz);
  }
  if (lg == 0) add_mom<LAYOUT>(p, i, plane, fx, fy, fz);
}

// --------------------------------------------------------------------------------------
// Gather on the column-major ELL list (transposed_list[i + k*pn]), one thread per row:
// the list read is coalesced across the warp.  cuda/kernel.cuh:102-134 and its ILP variants.
// --------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(1024)
lj_gather_ell(const void* __restrict__ q, void* __restrict__ p, int64_t pn, int64_t row_begin,
              int64_t row_end, int64_t plane, double c24, double c48, long long cl2_bits,
              const int32_t* __restrict__ tlist, const int32_t* __restrict__ nop) {
  const int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_end) return;
  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int32_t* __restrict__ col = tlist + i;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  int k = 0;
  for (; k + kUnroll <= np; k += kUnroll) {
    int j[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) j[u] = __ldg(col + (int64_t)(k + u) * pn);
    double xj[kUnroll], yj[kUnroll], zj[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) load_pos<LAYOUT>(q, j[u], plane, xj[u], yj[u], zj[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; u++)
      lj_pair(xj[u] - xi, yj[u] - yi, zj[u] - zi, c24, c48, cl2_bits, fx, fy, fz);
  }
  for (; k < np; k++) {
    const int j = __ldg(col + (int64_t)k * pn);
    double xj, yj, zj;
    load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
    lj_pair(xj - xi, yj - yi, zj - zi, c24, c48, cl2_bits, fx, fy, fz);
  }
  add_mom<LAYOUT>(p, i, plane, fx, fy, fz);
}

// --------------------------------------------------------------------------------------
// Newton's-third-law scatter on a half list (i<j): the i side accumulates in registers and is
// reduced by shuffles, the j side is scattered with RED.E.ADD.F64 (no return value).
// The *_with_aar kernels, cuda/kernel.cuh:238-469.  Summation order on p[j] is not
// deterministic -- compare with a tolerance only.
// --------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ __forceinline__ void red_mom(void* __restrict__ p, int64_t j, int64_t plane, double fx,
                                        double fy, double fz) {
  double* b;
  int64_t s;
  if (LAYOUT == LJ_AOS_D4) { b = reinterpret_cast<double*>(p) + 4 * j; s = 1; }
  else if (LAYOUT == LJ_AOS_D3) { b = reinterpret_cast<double*>(p) + 3 * j; s = 1; }
  else { b = reinterpret_cast<double*>(p) + j; s = plane; }
  atomicAdd(b, fx);
  atomicAdd(b + s, fy);
  atomicAdd(b + 2 * s, fz);
}

template <int G, int LAYOUT, bool PTR64>
__global__ void __launch_bounds__(1024)
lj_newton3_csr(const void* __restrict__ q, void* __restrict__ p, int64_t row_begin, int64_t row_end,
               int64_t plane, double c24, double c48, long long cl2_bits,
               const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
               const void* __restrict__ pointer) {
  const int rows_per_block = blockDim.x / G;
  const int64_t i = row_begin + (int64_t)blockIdx.x * rows_per_block + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  if (i >= row_end) return;
  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int32_t* __restrict__ row = list + row_offset<PTR64>(pointer, i);
  double fx = 0.0, fy = 0.0, fz = 0.0;
  for (int k = lg; k < np; k += G) {
    const int j = __ldg(row + k);
    double xj, yj, zj;
    load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    lj_pair(xj - xi, yj - yi, zj - zi, c24, c48, cl2_bits, gx, gy, gz);
    if (gx != 0.0 || gy != 0.0 || gz != 0.0) {
      fx += gx; fy += gy; fz += gz;
      red_mom<LAYOUT>(p, j, plane, -gx, -gy, -gz);
    }
  }
  if (G > 1) {
    fx = group_sum<G>(fx);
    fy = group_sum<G>(fy);
    fz = group_sum<G>(fz);
  }
  // p[i] also receives reactions from other rows concurrently -> atomic here too
  if (lg == 0) red_mom<LAYOUT>(p, i, plane, fx, fy, fz);
}

// -------------------------------------------------------------------------- dispatch ---
template <int G, int LAYOUT, bool PTR64>
void launch_csr(const lj_force_args* a, int64_t r0, int64_t r1, int tb, double c24, double c48,
                long long cl2_bits, bool newton3, cudaStream_t st) {
  const int rows_per_block = tb / G;
  const int64_t rows = r1 - r0;
  const unsigned blocks = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
  if (newton3)
    lj_newton3_csr<G, LAYOUT, PTR64><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                             c48, cl2_bits, a->list,
                                                             a->number_of_partners, a->pointer);
  else
    lj_gather_csr<G, LAYOUT, PTR64><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                            c48, cl2_bits, a->list,
                                                            a->number_of_partners, a->pointer);
}

template <int LAYOUT, bool PTR64>
bool launch_csr_g(int g, const lj_force_args* a, int64_t r0, int64_t r1, int tb, double c24,
                  double c48, long long cl2_bits, bool n3, cudaStream_t st) {
  switch (g) {
    case 1: launch_csr<1, LAYOUT, PTR64>(a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 2: launch_csr<2, LAYOUT, PTR64>(a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 4: launch_csr<4, LAYOUT, PTR64>(a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 8: launch_csr<8, LAYOUT, PTR64>(a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 16: launch_csr<16, LAYOUT, PTR64>(a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 32: launch_csr<32, LAYOUT, PTR64>(a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
  }
  return false;
}

template <int LAYOUT>
bool launch_layout(int g, const lj_force_args* a, int64_t r0, int64_t r1, int tb, double c24,
                   double c48, long long cl2_bits, bool n3, cudaStream_t st) {
  if (a->list_layout == LJ_LIST_ELL) {
    const int64_t rows = r1 - r0;
    const unsigned blocks = (unsigned)((rows + tb - 1) / tb);
    lj_gather_ell<LAYOUT><<<blocks, tb, 0, st>>>(a->q, a->p, a->pn, r0, r1, a->plane_stride, c24, c48,
                                                  cl2_bits, a->list, a->number_of_partners);
    return true;
  }
  return a->pointer64 ? launch_csr_g<LAYOUT, true>(g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st)
                      : launch_csr_g<LAYOUT, false>(g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st);
}

}  // namespace

int lj_force_tile_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, int g,
                         double c24, double c48, long long cl2_bits, cudaStream_t st);
int lj_force_mixed_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, int g, int tb,
                          cudaStream_t st);
bool lj_cluster_usable(const lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1);
int lj_force_cluster_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, double c24,
                            double c48, long long cl2_bits, cudaStream_t st);

int lj_force_launch(lj_ctx* ctx, const lj_force_args* a, cudaStream_t st) {
  LJ_REQUIRE(ctx, a != nullptr, "lj_force_step: null args");
  LJ_REQUIRE(ctx, a->pn >= 0, "lj_force_step: negative particle_number");
  if (a->pn == 0) return LJ_OK;
  LJ_REQUIRE(ctx, a->q && a->p && a->list && a->number_of_partners, "lj_force_step: null array");
  LJ_REQUIRE(ctx, a->list_layout == LJ_LIST_CSR || a->list_layout == LJ_LIST_ELL,
             "lj_force_step: unknown list layout");
  LJ_REQUIRE(ctx, a->list_layout == LJ_LIST_ELL || a->pointer != nullptr,
             "lj_force_step: CSR list needs pointer[]");
  LJ_REQUIRE(ctx, a->layout == LJ_AOS_D3 || a->layout == LJ_AOS_D4 || a->layout == LJ_SOA_D,
             "lj_force_step: layout must be AOS_D3, AOS_D4 or SOA_D");
  if (a->layout == LJ_AOS_D4)
    LJ_REQUIRE(ctx, ((uintptr_t)a->q % 32 == 0) && ((uintptr_t)a->p % 32 == 0),
               "lj_force_step: double4 arrays must be 32-byte aligned");
  if (a->layout == LJ_SOA_D)
    LJ_REQUIRE(ctx, a->plane_stride >= a->pn, "lj_force_step: SoA plane_stride < particle_number");
  int64_t r0 = a->row_begin, r1 = a->row_end;
  if (r0 == 0 && r1 == 0) r1 = a->pn;
  LJ_REQUIRE(ctx, 0 <= r0 && r0 <= r1 && r1 <= a->pn, "lj_force_step: bad row range");
  if (r0 == r1) return LJ_OK;

  int tb = a->threads_per_block ? a->threads_per_block : 128;
  // the reference CLI accepts 64..1024 (cuda/force_cuda.cu:380-383)
  LJ_REQUIRE(ctx, tb >= 64 && tb <= 1024 && tb % 32 == 0,
             "lj_force_step: THREAD_BLOCK must be a multiple of 32 in [64,1024]");

  int variant = a->variant;
  if (variant == LJ_VARIANT_AUTO || variant == LJ_VARIANT_CLUSTER) variant = LJ_VARIANT_SUBWARP;
  const bool n3 = variant == LJ_VARIANT_NEWTON3;
  LJ_REQUIRE(ctx, !(n3 && a->list_layout == LJ_LIST_ELL), "lj_force_step: Newton-3 needs a CSR list");
  int g = a->group;
  if (g == 0) g = (a->list_layout == LJ_LIST_ELL) ? 1 : 8;
  LJ_REQUIRE(ctx, g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32,
             "lj_force_step: group must be 1,2,4,8,16 or 32");

  const double c24 = 24.0 * a->dt, c48 = 48.0 * a->dt;
  long long cl2_bits;
  {
    double c = a->cl2;
    LJ_REQUIRE(ctx, c >= 0.0, "lj_force_step: negative CL2");
    memcpy(&cl2_bits, &c, sizeof c);
  }

  // AUTO prefers the cluster pair list when lj_build_list(LJ_LIST_CLUSTERS) mirrored exactly
  // these list arrays; LJ_VARIANT_CLUSTER insists on it
  if (a->variant == LJ_VARIANT_AUTO || a->variant == LJ_VARIANT_CLUSTER) {
    if (lj_cluster_usable(ctx, a, r0, r1)) return lj_force_cluster_launch(ctx, a, r0, r1, c24, c48, cl2_bits, st);
    LJ_REQUIRE(ctx, a->variant != LJ_VARIANT_CLUSTER,
               "lj_force_step: no cluster pair list for these arrays (build with LJ_LIST_CLUSTERS; FP64, "
               "CSR, row range on 4-row boundaries)");
  }
  if (a->precision == LJ_PREC_MIXED) {
    LJ_REQUIRE(ctx, !n3 && a->list_layout == LJ_LIST_CSR, "lj_force_step: mixed precision is gather/CSR only");
    return lj_force_mixed_launch(ctx, a, r0, r1, g, tb, st);
  }
  LJ_REQUIRE(ctx, a->precision == LJ_PREC_FP64, "lj_force_step: unknown precision");

  if (variant == LJ_VARIANT_TILE_TMA) {
    LJ_REQUIRE(ctx, a->list_layout == LJ_LIST_CSR, "lj_force_step: TILE_TMA needs a CSR list");
    return lj_force_tile_launch(ctx, a, r0, r1, g, c24, c48, cl2_bits, st);
  }
  LJ_REQUIRE(ctx, variant == LJ_VARIANT_SUBWARP || n3, "lj_force_step: unknown variant");

  bool ok = false;
  switch (a->layout) {
    case LJ_AOS_D3: ok = launch_layout<LJ_AOS_D3>(g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); break;
    case LJ_AOS_D4: ok = launch_layout<LJ_AOS_D4>(g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); break;
    case LJ_SOA_D: ok = launch_layout<LJ_SOA_D>(g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); break;
  }
  LJ_REQUIRE(ctx, ok, "lj_force_step: no kernel for this configuration");
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}
