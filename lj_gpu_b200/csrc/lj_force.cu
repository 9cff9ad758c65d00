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
#ifndef LJ_LIST_NO_ALLOCATE
#define LJ_LIST_NO_ALLOCATE 0
#endif
#ifndef LJ_LIST_PREFETCH_TRIPS
#define LJ_LIST_PREFETCH_TRIPS 0
#endif

// --------------------------------------------------------------------------------------
// Gather on a CSR list, G lanes per row.  G=32 is the reference's warp_unroll mapping
// (cuda/kernel.cuh:821-904), G=1 its thread-per-i mapping (cuda/kernel.cuh:67-100).
// --------------------------------------------------------------------------------------
// MODE 0: the kernel.  MODE 1/2 are diagnostics selected by variant 100/101 (tools/sweep.py):
// 1 = memory path only (pair math replaced by one add per component), 2 = math only (gather
// confined to 1024 L1-resident particles).  They bracket what the real kernel can reach.
template <int G, int LAYOUT, bool PTR64, int MODE = 0>
__global__ void __launch_bounds__(1024)
lj_gather_csr(const void* __restrict__ q, void* __restrict__ p, int64_t row_begin, int64_t row_end,
              int64_t plane, double c24, double c48, long long cl2_bits,
              const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
              const void* __restrict__ pointer, cudaTextureObject_t ltex = 0,
              const unsigned char* __restrict__ only_rows = nullptr) {
  const int rows_per_block = blockDim.x / G;
  const int64_t i = row_begin + (int64_t)blockIdx.x * rows_per_block + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  // whole groups leave together (G divides 32), so the shuffles below stay convergent
  if (i >= row_end) return;
  if (only_rows && !only_rows[i]) return;  // lj_list_mirror: only the rows the mirror does not hold

  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int64_t roff = row_offset<PTR64>(pointer, i);
  const int32_t* __restrict__ row = list + roff;
  // MODE 5 (experiment): list words through the TEX pipe instead of the LSU pipe
  auto ldl = [&](int kk) -> int {
    if (MODE == 5) return tex1Dfetch<int>(ltex, (int)(roff + kk));
#if LJ_LIST_NO_ALLOCATE
    int v;  // the list is used once: do not let it displace q lines from L1
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(row + kk));
    return v;
#else
    return __ldg(row + kk);
#endif
  };

  double fx = 0.0, fy = 0.0, fz = 0.0;
  int k = lg;
  // Full tiles: kUnroll independent gathers per lane and trip.  The list words of the NEXT trip
  // are requested before this trip's gathers: every trip of a group starts a fresh 128-byte line
  // of the list that comes straight from HBM (~1 us), and without the prefetch that latency sits
  // in front of every gather (ncu: long-scoreboard stalls on the first use of j).
  int jn[kUnroll];
  const bool any_full = k + (kUnroll - 1) * G < np;
  if (any_full) {
#pragma unroll
    for (int u = 0; u < kUnroll; u++) jn[u] = ldl(k + u * G);
  }
  for (; k + (kUnroll - 1) * G < np; k += kUnroll * G) {
    int j[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) j[u] = jn[u];
    if (k + kUnroll * G + (kUnroll - 1) * G < np) {
#pragma unroll
      for (int u = 0; u < kUnroll; u++) jn[u] = ldl(k + kUnroll * G + u * G);
    }
#if LJ_LIST_PREFETCH_TRIPS > 0
    // pull the list line that is LJ_LIST_PREFETCH_TRIPS trips ahead towards L2/L1 (no register,
    // no fault); one lane per group issues it
    if (lg == 0 && k + LJ_LIST_PREFETCH_TRIPS * kUnroll * G < np)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(row + k + LJ_LIST_PREFETCH_TRIPS * kUnroll * G));
#endif
    if (MODE == 2) {
#pragma unroll
      for (int u = 0; u < kUnroll; u++) j[u] &= 1023;
    }
    if (MODE == 3) {  // list only
#pragma unroll
      for (int u = 0; u < kUnroll; u++) fx += (double)j[u];
      continue;
    }
    if (MODE == 4) {  // gather only: indices made up from the loop counter, no list traffic
#pragma unroll
      for (int u = 0; u < kUnroll; u++) j[u] = (int)i + ((k + u * G) & 255) - 128 < 0 ? 0 : (int)i + ((k + u * G) & 255) - 128;
    }
    double xj[kUnroll], yj[kUnroll], zj[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      if (MODE == 6) {  // experiment: positions through the TEX pipe (two 16-byte fetches)
        const int4 a4 = tex1Dfetch<int4>(ltex, 2 * j[u]), b4 = tex1Dfetch<int4>(ltex, 2 * j[u] + 1);
        xj[u] = __hiloint2double(a4.y, a4.x); yj[u] = __hiloint2double(a4.w, a4.z);
        zj[u] = __hiloint2double(b4.y, b4.x);
      } else {
        load_pos<LAYOUT>(q, j[u], plane, xj[u], yj[u], zj[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      if (MODE == 1 || MODE == 4) { fx += xj[u]; fy += yj[u]; fz += zj[u]; }
      else lj_pair(xj[u] - xi, yj[u] - yi, zj[u] - zi, c24, c48, cl2_bits, fx, fy, fz);
    }
  }
  for (; k < np; k += G) {
    int j = ldl(k);
    if (MODE == 2) j &= 1023;
    double xj, yj, zj;
    load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
    if (MODE == 1) { fx += xj; fy += yj; fz += zj; }
    else lj_pair(xj - xi, yj - yi, zj - zi, c24, c48, cl2_bits, fx, fy, fz);
  }

  if (G > 1) {
    fx = group_sum<G>(fx);
    fy = group_sum<G>(fy);
    fz = group_sum<G>(fz);
  }
  if (lg == 0) add_mom<LAYOUT>(p, i, plane, fx, fy, fz);
}

// --------------------------------------------------------------------------------------
// List words fetched four at a time.  Diagnostics (tools/diag_force.py, config C): the scalar
// kernel is bound by the L1/LSU pipe -- "memory instructions only" takes 0.39 of its 0.42 ms,
// of which the 4-byte list loads are 0.17-0.22 ms (one LDG.32 per pair-iteration touching four
// sectors) and the 256-bit gathers 0.235 ms (16 SM-cycles per warp instruction whatever the
// address pattern).  Here every lane loads one 16-byte-ALIGNED int4 of its row (G lanes = one
// contiguous 16G-byte span) and uses its own four consecutive entries as the four unrolled
// gathers: one list instruction per FOUR pair-iterations.  Entries of the first/last int4 that
// lie outside the row are masked (cutoff -1).  Needs a 16-byte aligned list and a known
// allocation length so that the last int4 never reads past it.  Summation order differs from the
// scalar kernel (lane <-> entry mapping), results agree to rounding.
// --------------------------------------------------------------------------------------
template <int G, int LAYOUT, bool PTR64>
__global__ void __launch_bounds__(1024)
lj_gather_csr_v4(const void* __restrict__ q, void* __restrict__ p, int64_t row_begin, int64_t row_end,
                 int64_t plane, double c24, double c48, long long cl2_bits,
                 const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
                 const void* __restrict__ pointer, int64_t list_entries) {
  static_assert(G >= 4 && G % 4 == 0, "the int4 transpose needs G a multiple of 4");
  const int rows_per_block = blockDim.x / G;
  const int64_t i = row_begin + (int64_t)blockIdx.x * rows_per_block + threadIdx.x / G;
  const int lane = threadIdx.x & 31;
  const int lg = threadIdx.x % G;
  if (i >= row_end) return;  // whole groups leave together
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - lg));

  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int64_t off = row_offset<PTR64>(pointer, i);
  const int64_t a0 = off & ~(int64_t)3;      // aligned start of the row's first int4
  const int lead = (int)(off - a0);          // entries of that int4 belonging to the previous row
  const int total = lead + np;               // aligned entries up to the end of the row
  const int4* __restrict__ row4 = reinterpret_cast<const int4*>(list + a0);

  double fx = 0.0, fy = 0.0, fz = 0.0;
  for (int c0 = 0; c0 < total; c0 += 4 * G) {  // uniform within the group
    // my int4 of this chunk: aligned entries [c0 + 4 lg, c0 + 4 lg + 4)
    int4 v = make_int4((int)i, (int)i, (int)i, (int)i);
    const int e0 = c0 + 4 * lg;
    if (e0 < total) {
      if (a0 + e0 + 4 <= list_entries) {
        v = __ldg(row4 + (e0 >> 2));
      } else {  // the very last int4 of the array may be partial
        const int32_t* s = list + a0 + e0;
        if (a0 + e0 + 0 < list_entries) v.x = __ldg(s + 0);
        if (a0 + e0 + 1 < list_entries) v.y = __ldg(s + 1);
        if (a0 + e0 + 2 < list_entries) v.z = __ldg(s + 2);
      }
    }
    // each lane keeps ITS four consecutive entries: the scattered gather costs the L1 pipe the
    // same ~16 cycles whatever the address pattern (tools/diag_force.py), so no transpose
    int j[4] = {v.x, v.y, v.z, v.w};
    double xj[4], yj[4], zj[4];
    long long lim[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u;  // aligned entry index of j[u]
      const bool valid = e >= lead && e < total;
      if (!valid) j[u] = (int)i;      // masked: any valid particle
      lim[u] = valid ? cl2_bits : -1ll;
      load_pos<LAYOUT>(q, j[u], plane, xj[u], yj[u], zj[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
      lj_pair(xj[u] - xi, yj[u] - yi, zj[u] - zi, c24, c48, lim[u], fx, fy, fz);
  }
  fx = group_sum<G>(fx);
  fy = group_sum<G>(fy);
  fz = group_sum<G>(fz);
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

// --------------------------------------------------------------------------------------
// Newton-3 on the HALF column-major ELL table, one thread per i (memopt2/memopt3_with_aar,
// cuda/kernel.cuh:344-423): coalesced list reads, i side in registers, j side by RED.
// --------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(1024)
lj_newton3_ell(const void* __restrict__ q, void* __restrict__ p, int64_t pn, int64_t row_begin,
               int64_t row_end, int64_t plane, double c24, double c48, long long cl2_bits,
               const int32_t* __restrict__ tlist, const int32_t* __restrict__ nop) {
  const int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_end) return;
  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int32_t* __restrict__ col = tlist + i;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  for (int k = 0; k < np; k++) {
    const int j = __ldg(col + (int64_t)k * pn);
    double xj, yj, zj;
    load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    lj_pair(xj - xi, yj - yi, zj - zi, c24, c48, cl2_bits, gx, gy, gz);
    if (gx != 0.0 || gy != 0.0 || gz != 0.0) {
      fx += gx; fy += gy; fz += gz;
      red_mom<LAYOUT>(p, j, plane, -gx, -gy, -gz);
    }
  }
  // p[i] also receives reactions from other rows concurrently -> atomic here too
  red_mom<LAYOUT>(p, i, plane, fx, fy, fz);
}

// --------------------------------------------------------------------------------------
// Gather on the row-major padded table list[i*width + k] (LJ_LIST_ELL_ROWS, the corrected
// sorted_list2d of cuda/force_cuda.cu:242-253), G lanes per row: a row is one contiguous run, so
// the G lanes read consecutive words like on the CSR list, without pointer[].
// --------------------------------------------------------------------------------------
template <int G, int LAYOUT>
__global__ void __launch_bounds__(1024)
lj_gather_ellrows(const void* __restrict__ q, void* __restrict__ p, int64_t row_begin, int64_t row_end,
                  int64_t plane, double c24, double c48, long long cl2_bits,
                  const int32_t* __restrict__ list, const int32_t* __restrict__ nop, int64_t width) {
  const int rows_per_block = blockDim.x / G;
  const int64_t i = row_begin + (int64_t)blockIdx.x * rows_per_block + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  if (i >= row_end) return;  // whole groups leave together (G divides 32)
  double xi, yi, zi;
  load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
  const int np = __ldg(nop + i);
  const int32_t* __restrict__ row = list + i * width;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  int k = lg;
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

// -------------------------------------------------------------------------- dispatch ---
template <int G, int LAYOUT, bool PTR64>
void launch_csr(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, int tb, double c24, double c48,
                long long cl2_bits, bool newton3, cudaStream_t st) {
  const int rows_per_block = tb / G;
  const int64_t rows = r1 - r0;
  const unsigned blocks = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
  if (newton3)
    lj_newton3_csr<G, LAYOUT, PTR64><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                             c48, cl2_bits, a->list,
                                                             a->number_of_partners, a->pointer);
  else if (G >= 4 && a->list_entries > 0 && (uintptr_t)a->list % 16 == 0 && a->list_scalar == 2 && a->variant < 100)
    lj_gather_csr_v4<(G >= 4 ? G : 4), LAYOUT, PTR64><<<blocks, tb, 0, st>>>(
        a->q, a->p, r0, r1, a->plane_stride, c24, c48, cl2_bits, a->list, a->number_of_partners,
        a->pointer, a->list_entries);
#if LJ_DIAG  // kernel-surgery variants 100..105 (tools/diag_force.py): diagnostic builds only
  else if (a->variant == 100 && G == 8 && LAYOUT == LJ_AOS_D4)
    lj_gather_csr<G, LAYOUT, PTR64, 1><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                               c48, cl2_bits, a->list,
                                                               a->number_of_partners, a->pointer);
  else if (a->variant == 101 && G == 8 && LAYOUT == LJ_AOS_D4)
    lj_gather_csr<G, LAYOUT, PTR64, 2><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                               c48, cl2_bits, a->list,
                                                               a->number_of_partners, a->pointer);
  else if (a->variant == 102 && G == 8 && LAYOUT == LJ_AOS_D4)
    lj_gather_csr<G, LAYOUT, PTR64, 3><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                               c48, cl2_bits, a->list,
                                                               a->number_of_partners, a->pointer);
  else if (a->variant == 104 && G == 8 && LAYOUT == LJ_AOS_D4 && a->list_entries > 0 && a->list_entries < (1LL << 27)) {
    static cudaTextureObject_t tex = 0;
    static const void* tex_ptr = nullptr;
    if (tex_ptr != a->list) {
      if (tex) cudaDestroyTextureObject(tex);
      cudaResourceDesc rd{};
      rd.resType = cudaResourceTypeLinear;
      rd.res.linear.devPtr = const_cast<int32_t*>(a->list);
      rd.res.linear.desc = cudaCreateChannelDesc<int>();
      rd.res.linear.sizeInBytes = (size_t)a->list_entries * 4;
      cudaTextureDesc td{};
      td.readMode = cudaReadModeElementType;
      cudaCreateTextureObject(&tex, &rd, &td, nullptr);
      tex_ptr = a->list;
    }
    lj_gather_csr<G, LAYOUT, PTR64, 5><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                               c48, cl2_bits, a->list,
                                                               a->number_of_partners, a->pointer, tex);
  }
  else if (a->variant == 105 && G == 8 && LAYOUT == LJ_AOS_D4 && a->pn < (1LL << 26)) {
    static cudaTextureObject_t qtex = 0;
    static const void* qtex_ptr = nullptr;
    if (qtex_ptr != a->q) {
      if (qtex) cudaDestroyTextureObject(qtex);
      cudaResourceDesc rd{};
      rd.resType = cudaResourceTypeLinear;
      rd.res.linear.devPtr = const_cast<void*>(a->q);
      rd.res.linear.desc = cudaCreateChannelDesc<int4>();
      rd.res.linear.sizeInBytes = (size_t)a->pn * 32;
      cudaTextureDesc td{};
      td.readMode = cudaReadModeElementType;
      cudaCreateTextureObject(&qtex, &rd, &td, nullptr);
      qtex_ptr = a->q;
    }
    lj_gather_csr<G, LAYOUT, PTR64, 6><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                               c48, cl2_bits, a->list,
                                                               a->number_of_partners, a->pointer, qtex);
  }
  else if (a->variant == 103 && G == 8 && LAYOUT == LJ_AOS_D4)
    lj_gather_csr<G, LAYOUT, PTR64, 4><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                               c48, cl2_bits, a->list,
                                                               a->number_of_partners, a->pointer);
#endif
  else {
    // the gather kernels use no shared memory: give the whole 228 KB array to L1
    // (per context = per device: the attribute is a property of the function ON a device)
    int& carved = ctx->func_occ[reinterpret_cast<const void*>(lj_gather_csr<G, LAYOUT, PTR64>)];
    if (!carved) {
      cudaFuncSetAttribute(lj_gather_csr<G, LAYOUT, PTR64>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxL1);
      carved = 1;
    }
    lj_gather_csr<G, LAYOUT, PTR64><<<blocks, tb, 0, st>>>(a->q, a->p, r0, r1, a->plane_stride, c24,
                                                            c48, cl2_bits, a->list,
                                                            a->number_of_partners, a->pointer, 0, ctx->only_rows_launch);
  }
}

template <int LAYOUT, bool PTR64>
bool launch_csr_g(lj_ctx* ctx, int g, const lj_force_args* a, int64_t r0, int64_t r1, int tb, double c24,
                  double c48, long long cl2_bits, bool n3, cudaStream_t st) {
  switch (g) {
    case 1: launch_csr<1, LAYOUT, PTR64>(ctx, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 2: launch_csr<2, LAYOUT, PTR64>(ctx, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 4: launch_csr<4, LAYOUT, PTR64>(ctx, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 8: launch_csr<8, LAYOUT, PTR64>(ctx, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 16: launch_csr<16, LAYOUT, PTR64>(ctx, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
    case 32: launch_csr<32, LAYOUT, PTR64>(ctx, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); return true;
  }
  return false;
}

template <int LAYOUT>
bool launch_layout(lj_ctx* ctx, int g, const lj_force_args* a, int64_t r0, int64_t r1, int tb, double c24,
                   double c48, long long cl2_bits, bool n3, cudaStream_t st) {
  if (a->list_layout == LJ_LIST_ELL) {
    const int64_t rows = r1 - r0;
    const unsigned blocks = (unsigned)((rows + tb - 1) / tb);
    if (n3)
      lj_newton3_ell<LAYOUT><<<blocks, tb, 0, st>>>(a->q, a->p, a->pn, r0, r1, a->plane_stride, c24, c48,
                                                     cl2_bits, a->list, a->number_of_partners);
    else
      lj_gather_ell<LAYOUT><<<blocks, tb, 0, st>>>(a->q, a->p, a->pn, r0, r1, a->plane_stride, c24, c48,
                                                    cl2_bits, a->list, a->number_of_partners);
    return true;
  }
  if (a->list_layout == LJ_LIST_ELL_ROWS) {
    const int64_t rows = r1 - r0;
#define LJ_ELLROWS(GG)                                                                              \
    lj_gather_ellrows<GG, LAYOUT><<<(unsigned)((rows + tb / GG - 1) / (tb / GG)), tb, 0, st>>>(     \
        a->q, a->p, r0, r1, a->plane_stride, c24, c48, cl2_bits, a->list, a->number_of_partners, a->ell_width)
    switch (g) {
      case 1: LJ_ELLROWS(1); return true;
      case 2: LJ_ELLROWS(2); return true;
      case 4: LJ_ELLROWS(4); return true;
      case 8: LJ_ELLROWS(8); return true;
      case 16: LJ_ELLROWS(16); return true;
      case 32: LJ_ELLROWS(32); return true;
    }
#undef LJ_ELLROWS
    return false;
  }
  return a->pointer64 ? launch_csr_g<LAYOUT, true>(ctx, g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st)
                      : launch_csr_g<LAYOUT, false>(ctx, g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st);
}

}  // namespace

int lj_force_tile_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, int g,
                         double c24, double c48, long long cl2_bits, cudaStream_t st);
int lj_force_mixed_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, int g, int tb,
                          cudaStream_t st);
bool lj_cluster_usable(const lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1);
int lj_force_cluster_launch(lj_ctx* ctx, const lj_force_args* a, int64_t r0, int64_t r1, double c24,
                            double c48, long long cl2_bits, cudaStream_t st);

int lj_force_launch(lj_ctx* ctx, const lj_force_args* a, cudaStream_t st, int part) {
  LJ_REQUIRE(ctx, a != nullptr, "lj_force_step: null args");
  LJ_REQUIRE(ctx, a->pn >= 0, "lj_force_step: negative particle_number");
  if (a->pn == 0) return LJ_OK;
  LJ_REQUIRE(ctx, a->q && a->p && a->list && a->number_of_partners, "lj_force_step: null array");
  LJ_REQUIRE(ctx, a->list_layout == LJ_LIST_CSR || a->list_layout == LJ_LIST_ELL || a->list_layout == LJ_LIST_ELL_ROWS,
             "lj_force_step: unknown list layout");
  LJ_REQUIRE(ctx, a->list_layout != LJ_LIST_CSR || a->pointer != nullptr,
             "lj_force_step: CSR list needs pointer[]");
  LJ_REQUIRE(ctx, a->list_layout != LJ_LIST_ELL_ROWS || a->ell_width > 0,
             "lj_force_step: LJ_LIST_ELL_ROWS needs ell_width");
  LJ_REQUIRE(ctx, a->layout == LJ_AOS_D3 || a->layout == LJ_AOS_D4 || a->layout == LJ_SOA_D ||
                      ((a->layout == LJ_AOS_F4 || a->layout == LJ_AOS_F3) && a->precision == LJ_PREC_MIXED),
             "lj_force_step: layout must be AOS_D3, AOS_D4, SOA_D, or AOS_F4 / AOS_F3 with LJ_PREC_MIXED");
  if (a->layout == LJ_AOS_F3)
    LJ_REQUIRE(ctx, ((uintptr_t)a->q % 4 == 0) && ((uintptr_t)a->p % 4 == 0),
               "lj_force_step: float3 arrays must be 4-byte aligned");
  if (a->layout == LJ_AOS_F4)
    LJ_REQUIRE(ctx, ((uintptr_t)a->q % 16 == 0) && ((uintptr_t)a->p % 16 == 0),
               "lj_force_step: float4 arrays must be 16-byte aligned");
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
  if (variant == LJ_VARIANT_AUTO || variant == LJ_VARIANT_CLUSTER || variant >= 100) variant = LJ_VARIANT_SUBWARP;
  const bool n3 = variant == LJ_VARIANT_NEWTON3;
  LJ_REQUIRE(ctx, !(n3 && a->list_layout == LJ_LIST_ELL_ROWS), "lj_force_step: Newton-3 needs a CSR or ELL list");
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

  // LJ_VARIANT_CLUSTER runs on the cluster pair list that lj_build_list(LJ_LIST_CLUSTERS)
  // mirrored for exactly these list arrays.  AUTO stays on the per-row gather: on B200 the
  // cluster kernel trades L1 wavefronts for 1.5x FP64 work and measures slower.
  if (a->variant == LJ_VARIANT_CLUSTER) {
    LJ_REQUIRE(ctx, lj_cluster_usable(ctx, a, r0, r1),
               "lj_force_step: no cluster pair list for these arrays (build with LJ_LIST_CLUSTERS; "
               "CSR, row range on 4-row boundaries)");
    if (a->precision == LJ_PREC_MIXED) return lj_force_mixed_launch(ctx, a, r0, r1, g, tb, st);
    return lj_force_cluster_launch(ctx, a, r0, r1, c24, c48, cl2_bits, st);
  }
  // The cell-tile mirror (lj_build_list with LJ_LIST_TILES) serves FP64 and mixed calls on exactly the
  // arrays and row range it was built for; AUTO prefers it.
  if (a->variant == LJ_VARIANT_CELLTILE)
    LJ_REQUIRE(ctx, lj_celltile_usable(ctx, a, r0, r1),
               "lj_force_step: no cell-tile mirror for these arrays (build with LJ_LIST_TILES; CSR, "
               "double positions, same row range)");
  if (part != 0) {  // lj_force_step_part: INTERIOR / BOUNDARY tiles of the mirror
    LJ_REQUIRE(ctx, part == LJ_PART_INTERIOR || part == LJ_PART_BOUNDARY, "lj_force_step_part: unknown part");
    LJ_REQUIRE(ctx, (a->variant == LJ_VARIANT_CELLTILE || a->variant == LJ_VARIANT_AUTO) && lj_celltile_usable(ctx, a, r0, r1),
               "lj_force_step_part: needs the cell-tile mirror of these arrays and this row range (LJ_LIST_TILES)");
    return lj_force_celltile_launch(ctx, a, c24, c48, cl2_bits, st, part);
  }
  if ((a->variant == LJ_VARIANT_CELLTILE || (a->variant == LJ_VARIANT_AUTO && lj_celltile_worthwhile(ctx))) &&
      lj_celltile_usable(ctx, a, r0, r1))
  {
    int rc = lj_force_celltile_launch(ctx, a, c24, c48, cl2_bits, st, 0);
    if (rc || ctx->tl_outside == 0) return rc;
    // lj_list_mirror left some rows out (an entry outside their tile's region): the per-row kernel, those rows only
    ctx->only_rows_launch = ctx->tl_rowflag;
    bool ok = false;
    switch (a->layout) {
      case LJ_AOS_D3: ok = launch_layout<LJ_AOS_D3>(ctx, 8, a, r0, r1, tb, c24, c48, cl2_bits, false, st); break;
      case LJ_AOS_D4: ok = launch_layout<LJ_AOS_D4>(ctx, 8, a, r0, r1, tb, c24, c48, cl2_bits, false, st); break;
      case LJ_SOA_D: ok = launch_layout<LJ_SOA_D>(ctx, 8, a, r0, r1, tb, c24, c48, cl2_bits, false, st); break;
    }
    ctx->only_rows_launch = nullptr;
    LJ_REQUIRE(ctx, ok, "lj_force_step: no per-row kernel for the rows outside the mirror");
    LJ_LAUNCHED(ctx);
    return LJ_OK;
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
    case LJ_AOS_D3: ok = launch_layout<LJ_AOS_D3>(ctx, g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); break;
    case LJ_AOS_D4: ok = launch_layout<LJ_AOS_D4>(ctx, g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); break;
    case LJ_SOA_D: ok = launch_layout<LJ_SOA_D>(ctx, g, a, r0, r1, tb, c24, c48, cl2_bits, n3, st); break;
  }
  LJ_REQUIRE(ctx, ok, "lj_force_step: no kernel for this configuration");
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}
