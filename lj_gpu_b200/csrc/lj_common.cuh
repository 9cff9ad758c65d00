// lj_common.cuh -- context, error plumbing and device helpers shared by the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <unordered_map>

#include "../../include/lj_b200.h"

// ------------------------------------------------------------------------------------
// Context.  One per host thread / GPU; no global state (SURVEY 8b "Threading").
// ------------------------------------------------------------------------------------
struct lj_list_totals {      // written by the list-build kernels, read back on demand
  unsigned long long total;  // entries of the list
  int max_np;                // longest row
  int overflow;              // bit0: capacity, bit1: int32 pointer overflow, bit2: cluster list
  unsigned long long cl_total;  // entries of the cluster pair list
};

struct lj_grid_params {  // cell grid derived ON THE DEVICE from the bounding box
  double ox, oy, oz;     // origin (min corner)
  double inv_cell;       // 1 / cell edge
  int nx, ny, nz;
  int ncell;
};

// Geometry of the cell-tile mirror (lj_celltile.cuh): tiles of `tc` consecutive x-cells of one
// (y,z) pencil of the cell grid.  Written on the device by the list build, copied to the host.
struct lj_tile_geom {
  int tc;          // cells per tile along x
  int ntx;         // tiles per pencil
  int nx, ny, nz;  // cell grid
  int ntiles;      // ntx * ny * nz
  int max_rows;    // largest number of rows (particles) in one tile
  int max_yrow;    // longest y-row (five pencils), particles; cap_y = max_yrow + 8
  int max_units;   // largest list segment of one tile, in units of 8 entries
  int pad;         // the force kernel's unit counter lives here
  int ncols_active;  // columns (tx, cz) with at least one list entry; the others (ghost layers of a
  int pad2;          // decomposed run) are skipped by the force kernel.  pad2: columns selected by a part launch
  unsigned long long total_units;  // whole mirror list, units of 8 entries
};

struct lj_ctx {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 227 * 1024;  // cudaDevAttrMaxSharedMemoryPerBlockOptin of the device
  cudaStream_t stream = nullptr;       // the context's own stream
  cudaStream_t copy_stream = nullptr;  // staging-ring DMA
  cudaMemPool_t pool = nullptr;
  int64_t launches = 0;
  std::string err;

  // lj_kernel_timing: CUDA events around every launch of the dominant force kernel (lj_celltile_force), on the
  // launching stream; a ring of event pairs, folded into the sums when a pair is reused or the sums are read
  bool kt_on = false;
  cudaEvent_t kt_ev[2 * 64] = {};
  int kt_head = 0, kt_pending = 0;   // next pair to use; pairs recorded and not folded yet
  double kt_ms = 0.0;
  int64_t kt_launches = 0;

  // pinned staging ring for pageable host memory (lj_upload / lj_download)
  void* ring[2] = {nullptr, nullptr};
  cudaEvent_t ring_ev[2] = {nullptr, nullptr};
  size_t ring_bytes = 0;

  // list-build scratch, grow-only
  int64_t scratch_pn = 0;
  int64_t scratch_cells = 0;
  double* bbox = nullptr;            // 6 ordered-encoded doubles (as uint64)
  lj_grid_params* grid = nullptr;    // device
  int32_t* cell_of = nullptr;        // [pn] cell id of each particle
  int32_t* cell_slot = nullptr;      // [pn] arrival slot inside its cell
  uint32_t* cell_count = nullptr;    // [cells+1]
  uint32_t* cell_start = nullptr;    // [cells+1]
  double4* sorted_pos = nullptr;     // [pn] positions in cell order, .w = original index bits
  int32_t* sorted_tmp = nullptr;     // [pn] particle ids in arrival order
  unsigned long long* scan_tmp = nullptr;  // block sums for the scans
  int64_t scan_tmp_len = 0;
  lj_list_totals* totals = nullptr;  // device
  lj_list_totals* totals_host = nullptr;  // pinned
  int64_t last_capacity = 0;

  // cluster pair list: library-owned mirror of the CSR list most recently built with
  // LJ_LIST_CLUSTERS (union of 4 consecutive rows, entries (member mask << 28) | j)
  uint32_t* cl_list = nullptr;
  int64_t cl_cap = 0;
  long long* cl_ptr = nullptr;   // [nc+1]
  uint32_t* cl_cnt = nullptr;    // [nc+1]
  int64_t cl_nc_cap = 0;
  bool cl_valid = false;
  const void* cl_id_list = nullptr;  // identity of the CSR arrays it mirrors
  const void* cl_id_nop = nullptr;
  const void* cl_id_ptr = nullptr;
  int64_t cl_pn = 0, cl_r0 = 0, cl_r1 = 0, cl_entries = 0;

  // cell-tile mirror: library-owned copy of the CSR list most recently built with LJ_LIST_TILES,
  // rows in cell order, entries = 16-bit indices into the tile's staged region (lj_celltile.cuh)
  lj_tile_geom* tl_geom = nullptr;       // device
  lj_tile_geom* tl_geom_host = nullptr;  // pinned
  int32_t* tl_order = nullptr;           // [pn]   original index of sorted slot s
  int32_t* tl_cnt = nullptr;             // [pn]   entries of row s
  uint32_t* tl_units = nullptr;          // [pn+1] padded row length, units of 8 entries
  uint32_t* tl_off = nullptr;            // [pn+1] exclusive scan of tl_units
  double* tl_qs = nullptr;               // positions in cell order, refreshed every step: [cap+2] {x,y} ...
  double* tl_qz = nullptr;               // ... followed by [cap+2] z (inside the tl_qs allocation)
  double* soa6_q = nullptr;              // six-array SoA entry points: gathered q / p planes
  double* soa6_p = nullptr;
  int64_t soa6_cap = 0;
  int4* tl_qfx = nullptr;                // [pn+2] the same in 32-bit fixed point + original index (mixed precision)
  int64_t tl_qfx_cap = 0;
  uint32_t* tl_cell_start = nullptr;     // [ncell+1] private copy of the cell offsets
  uint16_t* tl_list = nullptr;
  uint2* tl_tab = nullptr;               // [ntiles][6] y-row table (lj_celltile.cuh)
  uint4* tl_ttab = nullptr;              // [ntiles][2] tile table
  int32_t* tl_cols = nullptr;            // [ncols_active] the active columns, ascending
  int64_t tl_cols_cap = 0;
  int32_t* tl_zflag = nullptr;           // [nz] cell layers holding a particle outside the row range (part launches)
  int32_t* tl_cols_sel = nullptr;        // columns selected by the last part launch (count in tl_geom->pad2)
  int64_t tl_sel_cap = 0;
  int4* tl_meta = nullptr;               // [pn] {entries, first unit, original index, 0} of row s
  int64_t tl_pn_cap = 0, tl_cells_cap = 0, tl_list_cap = 0, tl_tab_cap = 0;  // particles, cells, units, tiles
  bool tl_valid = false;
  const void* tl_id_list = nullptr;
  const void* tl_id_nop = nullptr;
  const void* tl_id_ptr = nullptr;
  int64_t tl_pn = 0, tl_r0 = 0, tl_r1 = 0;
  uint64_t tl_token = 0;                 // generation of the valid mirror (lj_list_mirror_token); callers pass it back
  uint64_t tl_token_next = 1;
  unsigned char* tl_rowflag = nullptr;   // lj_list_mirror: rows the mirror does not hold (an entry outside the tile's region)
  int64_t tl_rowflag_cap = 0;
  const unsigned char* only_rows_launch = nullptr;  // transient: row filter of the per-row launch in progress
  int64_t tl_outside = 0;                // how many such rows; they take the per-row kernel after the cell-tile kernel
  int32_t* tl_slot_of = nullptr;         // lj_list_mirror: cell-order slot of particle i
  int64_t tl_slot_cap = 0;
  lj_tile_geom tl_g{};                   // host copy of the geometry of the valid mirror

  // mixed-precision scratch: origin-shifted float4 positions
  float4* q32 = nullptr;
  int64_t q32_len = 0;

  // per-context (= per-device) record of what cudaFuncSetAttribute has been applied to which kernel:
  // the attributes are per device, a process-wide flag would leave a second GPU unconfigured
  std::unordered_map<const void*, size_t> func_smem;  // kernel -> opted-in dynamic shared memory
  std::unordered_map<const void*, int> func_occ;      // kernel -> resident CTAs per SM (occupancy query)

  // lj_measure: the tile engine allocates sorted_list itself right after its count pass (no sizing build)
  cudaEvent_t ev_copy = nullptr;  // lj_measure: copy-stream upload of p <-> the caller-visible stream
  int32_t* alloc_list = nullptr;
  int64_t alloc_capacity = 0;
  unsigned int* pull_counter = nullptr;  // lj_halo_pull_sync: blocks of the running copy that have finished
  long long* diag_buf = nullptr;  // LJ_DIAG builds only: per-warp cycle counters of the cell-tile kernel
  int diag_dumps = 0;

  // cached CUDA graph for lj_force_loop
  cudaGraphExec_t graph_exec = nullptr;
  lj_force_args graph_args{};
  int graph_loop = 0;
  int64_t graph_step_launches = 0;
};

int lj_set_error(lj_ctx* ctx, int status, const char* what, const char* detail);
cudaEvent_t lj_kernel_timing_event(lj_ctx* ctx, cudaStream_t st, bool first);  // lj_runtime.cu (lj_kernel_timing)

#define LJ_CUDA(ctx, call)                                                            \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) return lj_set_error((ctx), LJ_ERR_CUDA, #call, cudaGetErrorString(e__)); \
  } while (0)

#define LJ_REQUIRE(ctx, cond, msg)                                         \
  do {                                                                     \
    if (!(cond)) return lj_set_error((ctx), LJ_ERR_BAD_ARG, msg, #cond);   \
  } while (0)

// every kernel launch goes through this so gpu_launches can be reported truthfully
#define LJ_LAUNCHED(ctx)                                                                  \
  do {                                                                                    \
    (ctx)->launches++;                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) return lj_set_error((ctx), LJ_ERR_CUDA, "kernel launch", cudaGetErrorString(e__)); \
  } while (0)

// Diagnostic knobs (environment variables LJ_TILE_*) exist only in builds with -DLJ_DIAG=1
// (make DIAG=1); the product build reads no environment and carries no tuning side doors.
#ifndef LJ_DIAG
#define LJ_DIAG 0
#endif
static inline int lj_diag_int(const char* name) {
#if LJ_DIAG
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}
static inline bool lj_diag_set(const char* name) {
#if LJ_DIAG
  return getenv(name) != nullptr;
#else
  (void)name;
  return false;
#endif
}

// opt a kernel in to `bytes` of dynamic shared memory on this context's device (idempotent, cached per ctx).
// The attribute belongs to the DEVICE, not to the context: two contexts on one device (lj_decomp_args.devices
// may repeat an ordinal) would lower each other's setting if each asked for exactly what it needs, so a kernel
// is always opted in to the device's maximum; `bytes` above that fails here rather than at the launch.
#define LJ_FUNC_SMEM(ctx, kern, bytes)                                                              \
  do {                                                                                              \
    int rc__ = lj_func_smem((ctx), reinterpret_cast<const void*>(kern), (size_t)(bytes));           \
    if (rc__) return rc__;                                                                          \
  } while (0)
int lj_func_smem(lj_ctx* ctx, const void* kern, size_t bytes);  // lj_runtime.cu

// Top of every public entry point: a context belongs to ONE device; make it current for the calling
// host thread (a caller driving several GPUs from one thread switches contexts, not devices).
#define LJ_ENTER(ctx)                                                          \
  do {                                                                         \
    if (!(ctx)) return LJ_ERR_BAD_ARG;                                         \
    int cur__ = -1;                                                            \
    if (cudaGetDevice(&cur__) != cudaSuccess || cur__ != (ctx)->device)        \
      LJ_CUDA((ctx), cudaSetDevice((ctx)->device));                            \
  } while (0)

// (internal) lj_list_args.flags: sorted_list = NULL, capacity = 0 -> the tile engine allocates the list from the
// context's pool once its count pass knows the total and leaves it in ctx->alloc_list / alloc_capacity
constexpr int LJ_LIST_ALLOC_INTERNAL = 1 << 30;

// NULL is the CUDA legacy default stream, exactly as for a kernel launch: the reference runs
// everything there (cuda/force_cuda.cu:334) and torch hands out 0 for its default stream.
static inline cudaStream_t lj_stream(lj_ctx* ctx, void* s) { (void)ctx; return (cudaStream_t)s; }

// internal entry points shared between translation units
int lj_scratch_reserve(lj_ctx* ctx, int64_t pn, cudaStream_t st);
int lj_force_launch(lj_ctx* ctx, const lj_force_args* a, cudaStream_t st, int part = 0);
int lj_bbox_launch(lj_ctx* ctx, const void* q, int layout, int64_t pn, int64_t plane,
                   lj_list_totals* reset_totals, cudaStream_t st);
// cell-tile mirror (lj_nlist.cu builds it, lj_force_celltile.cu consumes it)
int lj_celltile_permute(lj_ctx* ctx, const lj_force_args* a, cudaStream_t st, int part);
bool lj_celltile_usable(const lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1);
bool lj_celltile_worthwhile(const lj_ctx* ctx);
int lj_force_celltile_launch(lj_ctx* ctx, const lj_force_args* a, double c24, double c48,
                             long long cl2_bits, cudaStream_t st, int part);
// shared memory of the force kernel: a ring of y-row slots (cap_y packed double3 records each) and
// a ring of list slots (a tile's list segment + 16 B of metadata per row)
constexpr size_t kTileSmemBudget = 227 * 1024 - 8 * 1024;  // minus the kernel's static shared memory
static inline int lj_celltile_cap_y(const lj_tile_geom& g) { return (g.max_yrow + 8 + 1) & ~1; }
static inline size_t lj_celltile_yslot_bytes(const lj_tile_geom& g) { return (size_t)lj_celltile_cap_y(g) * 24; }
static inline size_t lj_celltile_lslot_bytes(const lj_tile_geom& g) {
  return (size_t)g.max_units * 16 + (size_t)g.max_rows * 16;
}
constexpr int kTileMinYSlots = 7, kTileMinLSlots = 2;  // five rows in use + two in flight; two lists

// Fixed-point frame of the mixed-precision cell-tile kernel.  Coordinates are stored modulo 2^32
// counts of `unit`, a power of two chosen from the cutoff alone so that every difference the force
// needs is below 2^24 counts (its int -> float conversion is exact): cutoff 3.0 -> unit = 2^-22.
// No bounding box, nothing to clamp, and the precision does not depend on the size of the system.
// margin bounds |r2_f32 - r2_f64| near the cutoff: two half-count roundings per component
// (delta) and the FP32 roundings of the three-term sum; doubled.
struct lj_fx_frame { double scale, unit; float margin; };
static inline lj_fx_frame lj_fx_frame_for(double cl2) {
  const double c = sqrt(cl2);
  int e = 0;
  frexp(c, &e);  // c < 2^e
  lj_fx_frame f;
  f.unit = ldexp(1.0, e - 24);
  f.scale = ldexp(1.0, 24 - e);
  const double u = 5.9604644775390625e-8, dmax = 1.01 * c;
  const double delta = f.unit + u * dmax;
  f.margin = (float)(2.0 * (3.0 * (2.0 * dmax * delta + delta * delta) + 6.0 * u * dmax * dmax));
  return f;
}

// ------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------
#ifdef __CUDACC__

#ifndef LJ_POS_LOAD
#define LJ_POS_LOAD 256  // how a double4 position is fetched: 256 | 128 | 64
#endif

// order-preserving double <-> uint64 encoding (atomicMin/Max on doubles)
__device__ __forceinline__ unsigned long long enc_ordered(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_ordered(unsigned long long u) {
  u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}

// 256-bit read-only load of one double4 (one 32 B sector): LDG.E.256 on sm_100a.
__device__ __forceinline__ double4 ld_nc_d4(const double4* ptr) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(ptr));
  return v;
}
__device__ __forceinline__ double4 ld_d4(const double4* ptr) {
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(ptr)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_d4(double4* ptr, double4 v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "d"(v.x), "d"(v.y), "d"(v.z),
               "d"(v.w)
               : "memory");
}

// Position fetch for the three FP64 layouts.
template <int LAYOUT>
__device__ __forceinline__ void load_pos(const void* __restrict__ q, int64_t i, int64_t plane,
                                         double& x, double& y, double& z) {
  if (LAYOUT == LJ_AOS_D4) {
#if LJ_POS_LOAD == 256
    const double4 v = ld_nc_d4(reinterpret_cast<const double4*>(q) + i);
    x = v.x; y = v.y; z = v.z;
#elif LJ_POS_LOAD == 128   // LDG.128 (x,y) + LDG.64 (z): 24 of the 32 bytes
    const double2* b = reinterpret_cast<const double2*>(q) + 2 * i;
    const double2 xy = __ldg(b);
    x = xy.x; y = xy.y; z = __ldg(reinterpret_cast<const double*>(b + 1));
#else                      // three LDG.64
    const double* b = reinterpret_cast<const double*>(q) + 4 * i;
    x = __ldg(b); y = __ldg(b + 1); z = __ldg(b + 2);
#endif
  } else if (LAYOUT == LJ_AOS_D3) {
    const double* b = reinterpret_cast<const double*>(q) + 3 * i;
    x = __ldg(b); y = __ldg(b + 1); z = __ldg(b + 2);
  } else if (LAYOUT == LJ_AOS_F4) {  // float4 positions (cuda/force_cuda.cu:25): widened exactly
    const float4 v = __ldg(reinterpret_cast<const float4*>(q) + i);
    x = (double)v.x; y = (double)v.y; z = (double)v.z;
  } else if (LAYOUT == LJ_AOS_F3) {  // float3 positions (cuda/force_cuda.cu:24): widened exactly
    const float* b = reinterpret_cast<const float*>(q) + 3 * i;
    x = (double)__ldg(b); y = (double)__ldg(b + 1); z = (double)__ldg(b + 2);
  } else {
    const double* b = reinterpret_cast<const double*>(q) + i;
    x = __ldg(b); y = __ldg(b + plane); z = __ldg(b + 2 * plane);
  }
}

// p[i] += (fx,fy,fz); double4 keeps .w (read-modify-write of the whole 32 B vector, which
// is what the reference's `p[tid] = pf` does, cuda/kernel.cuh:118,133).
template <int LAYOUT>
__device__ __forceinline__ void add_mom(void* __restrict__ p, int64_t i, int64_t plane, double fx,
                                        double fy, double fz) {
  if (LAYOUT == LJ_AOS_D4) {
    double4* b = reinterpret_cast<double4*>(p) + i;
    double4 v = ld_d4(b);
    v.x += fx; v.y += fy; v.z += fz;
    st_d4(b, v);
  } else if (LAYOUT == LJ_AOS_D3) {
    double* b = reinterpret_cast<double*>(p) + 3 * i;
    b[0] += fx; b[1] += fy; b[2] += fz;
  } else if (LAYOUT == LJ_AOS_F4) {  // float4 momenta: one rounding per step, .w kept
    float4* b = reinterpret_cast<float4*>(p) + i;
    float4 v = *b;
    v.x = (float)((double)v.x + fx); v.y = (float)((double)v.y + fy); v.z = (float)((double)v.z + fz);
    *b = v;
  } else if (LAYOUT == LJ_AOS_F3) {  // float3 momenta: one rounding per step
    float* b = reinterpret_cast<float*>(p) + 3 * i;
    b[0] = (float)((double)b[0] + fx); b[1] = (float)((double)b[1] + fy); b[2] = (float)((double)b[2] + fz);
  } else {
    double* b = reinterpret_cast<double*>(p) + i;
    b[0] += fx; b[plane] += fy; b[2 * plane] += fz;
  }
}

// 1/a to ~1 ulp for normal positive a without the IEEE-division slow path:
// MUFU.RCP64H seed (rel. err < 2^-20) + one cubically convergent correction (3 DFMA).
__device__ __forceinline__ double fast_rcp(double a) {
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a));
  const double e = fma(-a, x0, 1.0);
  const double e2 = fma(e, e, e);
  return fma(x0, e2, x0);
}

// The FP64 pair body.  c24 = 24*dt, c48 = 48*dt (dt folded into the constants as
// cpu_ref/force_soa.cpp:202-203 does).  df = (24 r^6 - 48) / r^14 * dt = x^4 (24dt - 48dt x^3)
// with x = 1/r^2.  The cutoff test compares the bit patterns (positive doubles order like
// integers), which keeps it off the FP64 pipe; pairs with r2 > cl2 contribute nothing,
// r2 == cl2 contributes (cuda/kernel.cuh:30).
__device__ __forceinline__ void lj_pair(double dx, double dy, double dz, double c24, double c48,
                                        long long cl2_bits, double& fx, double& fy,
                                        double& fz) {
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
#if LJ_PAIR_FSEL
  // Mask the scalar, not the three products: ptxas turns a guarded accumulation (also one written
  // as @p fma in PTX) into six FSELs per pair, masking df costs two (the reference's "ifless"
  // form, kernel.cuh:59).
  const double x = fast_rcp(r2);
  const double x3 = x * x * x;
  const double t = fma(-c48, x3, c24);
  const double df = (__double_as_longlong(r2) <= cl2_bits) ? (x * x3) * t : 0.0;
#else
  // Mask the SEED of the reciprocal: MUFU.RCP64H delivers only a high word (the low word of
  // rcp.approx.ftz.f64 is zero by definition), so one 32-bit select makes x0 = +0 exactly, and then
  // e = 1, x = fma(0, 2, 0) = 0, x^3 = 0, df = 0 * c24 = 0: a pair beyond the cutoff adds exactly
  // zero, at the cost of one SEL instead of two FSEL on df.
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(r2));
  x0 = __hiloint2double((__double_as_longlong(r2) <= cl2_bits) ? __double2hiint(x0) : 0, 0);
  const double e = fma(-r2, x0, 1.0);
  const double x = fma(x0, fma(e, e, e), x0);
  const double x3 = x * x * x;
  const double df = (x * x3) * fma(-c48, x3, c24);
#endif
  fx = fma(df, dx, fx);
  fy = fma(df, dy, fy);
  fz = fma(df, dz, fz);
}

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

template <bool PTR64>
__device__ __forceinline__ int64_t row_offset(const void* __restrict__ pointer, int64_t i) {
  if (PTR64) return __ldg(reinterpret_cast<const long long*>(pointer) + i);
  // int32 pointer[] of the reference; offsets above 2^31-1 cannot be represented there
  return (int64_t)(uint32_t)__ldg(reinterpret_cast<const int*>(pointer) + i);
}

#endif  // __CUDACC__
