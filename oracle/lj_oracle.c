/*
 * lj_oracle.c -- CPU restatement of the lj_gpu hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA product path.  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py
 * may load it.  Nothing under lj_gpu_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every function below against
 *   - the reference's golden vectors ref_data/density0.5.dat and density1.dat
 *     (committed as tests/golden/density*.dat by tools/make_golden.py), and
 *   - full arrays (q, half list, p after 100 steps) produced by the real reference
 *     cpu_ref/force_soa.cpp compiled from /root/reference into oracle/_ref/ (see Makefile).
 *
 * What is restated (citations are relative to the reference repository root):
 *   ljo_init_fcc        cpu_ref/force_soa.cpp:42-50,115-140 ; cuda/force_cuda.cu:47-94
 *   ljo_makepair_brute  cpu_ref/force_soa.cpp:73-113 (half) ; cuda/force_cuda.cu:102-163 (full)
 *   ljo_makepair_cell   same output contract as ljo_makepair_brute, O(N) cell binning
 *                       (no counterpart in the reference: it only has the O(N^2) loop;
 *                       validated against ljo_makepair_brute and the real reference)
 *   ljo_force_sorted    cpu_ref/force_soa.cpp:163-195 (half list, Newton's 3rd law)
 *   ljo_force_gather    cuda/kernel.cuh:36-65 (full list, branch-free mask, p[i] only)
 *   ljo_shuffle_rows    cuda/force_cuda.cu:255-263 (per-row std::shuffle, mt19937(10))
 *   ljo_transpose_list  cuda/force_cuda.cu:229-240 (column-major ELL, zero padded)
 *
 * Numerics: build with -ffp-contract=off.  The only fused operation is the explicit
 * fma() chain in r2_search(), which fixes ONE contraction pattern for the list-membership
 * decision (dx*dx, then fma(dy,dy,.), then fma(dz,dz,.)); the CUDA list build uses the
 * identical chain, so membership is bit-exact by construction.  Force arithmetic follows
 * the reference expression order with every operation rounded separately.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------
 * mt19937 (32-bit Mersenne twister, the std::mt19937 parameter set) and the libstdc++
 * recipe for uniform_real_distribution<double>: generate_canonical<double,53> draws two
 * 32-bit words, lo first, and returns (lo + hi*2^32) / 2^64, clamped below 1.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  uint32_t s[624];
  int pos;
} ljo_mt;

static void mt_seed(ljo_mt *g, uint32_t seed) {
  g->s[0] = seed;
  for (int k = 1; k < 624; k++) {
    uint32_t prev = g->s[k - 1];
    g->s[k] = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)k;
  }
  g->pos = 624;
}

static void mt_refill(ljo_mt *g) {
  for (int k = 0; k < 624; k++) {
    uint32_t y = (g->s[k] & 0x80000000u) | (g->s[(k + 1) % 624] & 0x7fffffffu);
    uint32_t v = g->s[(k + 397) % 624] ^ (y >> 1);
    if (y & 1u) v ^= 0x9908b0dfu;
    g->s[k] = v;
  }
  g->pos = 0;
}

static uint32_t mt_next(ljo_mt *g) {
  if (g->pos >= 624) mt_refill(g);
  uint32_t y = g->s[g->pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static double mt_canonical(ljo_mt *g) {
  const double two32 = 4294967296.0;
  double lo = (double)mt_next(g);
  double hi = (double)mt_next(g);
  double r = (lo + hi * two32) / (two32 * two32);
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

/* uniform_real_distribution<double>(a, b)(mt) == canonical * (b - a) + a */
static double mt_uniform(ljo_mt *g, double a, double b) { return mt_canonical(g) * (b - a) + a; }

/* uniform_int_distribution<size_t>(0, hi)(mt) as libstdc++ (GCC >= 11) does it for a
 * 32-bit engine: when the requested range fits in 32 bits it uses Lemire's nearly
 * divisionless rejection on one 32-bit draw.  Needed for std::shuffle (ljo_shuffle_rows). */
static uint32_t mt_below(ljo_mt *g, uint32_t range /* returns value in [0, range) */) {
  uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (uint32_t)(-range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32);
}

/* ------------------------------------------------------------------------------------
 * System generator.  q is written as packed xyz triples (AoS3).  Returns the number of
 * particles, or -(needed) if cap is too small.  cells_out receives cells per side.
 * Order: iz outermost, then iy, ix; four basis atoms per cell; x, y, z jitter drawn in
 * that order from ONE mt19937(2) stream.
 * ---------------------------------------------------------------------------------- */
int64_t ljo_init_fcc(double density, double L, double *q_xyz, int64_t cap, int *cells_out) {
  const double s = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  const double hs = s * 0.5;
  const int n = (int)(L / s);
  if (cells_out) *cells_out = n;
  const int64_t need = 4LL * n * n * n;
  if (need > cap) return -need;
  ljo_mt g;
  mt_seed(&g, 2u);
  static const int basis[4][3] = {{0, 0, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
  int64_t pn = 0;
  for (int iz = 0; iz < n; iz++)
    for (int iy = 0; iy < n; iy++)
      for (int ix = 0; ix < n; ix++) {
        const double x = ix * s, y = iy * s, z = iz * s;
        for (int b = 0; b < 4; b++) {
          const double bx = basis[b][0] ? x + hs : x;
          const double by = basis[b][1] ? y + hs : y;
          const double bz = basis[b][2] ? z + hs : z;
          q_xyz[3 * pn + 0] = bx + mt_uniform(&g, 0.0, 0.1);
          q_xyz[3 * pn + 1] = by + mt_uniform(&g, 0.0, 0.1);
          q_xyz[3 * pn + 2] = bz + mt_uniform(&g, 0.0, 0.1);
          pn++;
        }
      }
  return pn;
}

/* The one contraction pattern used for the list-membership test, on CPU and GPU. */
static inline double r2_search(double dx, double dy, double dz) {
  return fma(dz, dz, fma(dy, dy, dx * dx));
}

/* ------------------------------------------------------------------------------------
 * Brute-force Verlet list, O(N^2).  full=1: every ordered (i,j), i!=j, r2 < sl2
 * (cuda/force_cuda.cu:138-144); full=0: i<j only (cpu_ref/force_soa.cpp:101-111).
 * Outputs the reference's three arrays: number_of_partners[pn], pointer[pn] (exclusive
 * scan, no sentinel), sorted_list[] with rows in ascending j.  Returns the number of list
 * entries, or -(needed) when cap is too small (nothing is written past cap).
 * ---------------------------------------------------------------------------------- */
int64_t ljo_makepair_brute(const double *q_xyz, int64_t pn, double sl2, int full,
                           int32_t *number_of_partners, int64_t *pointer,
                           int32_t *sorted_list, int64_t cap) {
  int64_t total = 0;
  for (int64_t i = 0; i < pn; i++) {
    const double xi = q_xyz[3 * i], yi = q_xyz[3 * i + 1], zi = q_xyz[3 * i + 2];
    int32_t cnt = 0;
    pointer[i] = total;
    for (int64_t j = full ? 0 : i + 1; j < pn; j++) {
      if (j == i) continue;
      const double dx = xi - q_xyz[3 * j], dy = yi - q_xyz[3 * j + 1], dz = zi - q_xyz[3 * j + 2];
      if (r2_search(dx, dy, dz) < sl2) {
        if (total + cnt < cap) sorted_list[total + cnt] = (int32_t)j;
        cnt++;
      }
    }
    number_of_partners[i] = cnt;
    total += cnt;
  }
  return total <= cap ? total : -total;
}

/* ------------------------------------------------------------------------------------
 * Brute-force rows for a SAMPLE of particles (systems too large for the O(N^2) build):
 * row k = every j != rows[k] with r2 < sl2 (full) or additionally j > rows[k] (half), ascending
 * in j -- the reference's membership test and row order (cuda/force_cuda.cu:122-145) applied
 * to the sampled i only, each against ALL pn particles.  out_ptr[k] is the offset of row k in
 * out_list (exclusive scan, nrows + 1 entries).  Returns the total, or -(needed) when cap is
 * too small.  Rows are independent: threaded.
 * ---------------------------------------------------------------------------------- */
int64_t ljo_rows_brute(const double *q_xyz, int64_t pn, double sl2, int full, const int64_t *rows,
                       int64_t nrows, int32_t *out_nop, int64_t *out_ptr, int32_t *out_list,
                       int64_t cap) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t k = 0; k < nrows; k++) {
    const int64_t i = rows[k];
    const double xi = q_xyz[3 * i], yi = q_xyz[3 * i + 1], zi = q_xyz[3 * i + 2];
    int32_t cnt = 0;
    for (int64_t j = full ? 0 : i + 1; j < pn; j++) {
      if (j == i) continue;
      const double dx = xi - q_xyz[3 * j], dy = yi - q_xyz[3 * j + 1], dz = zi - q_xyz[3 * j + 2];
      if (r2_search(dx, dy, dz) < sl2) cnt++;
    }
    out_nop[k] = cnt;
  }
  int64_t total = 0;
  for (int64_t k = 0; k < nrows; k++) { out_ptr[k] = total; total += out_nop[k]; }
  out_ptr[nrows] = total;
  if (total > cap) return -total;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t k = 0; k < nrows; k++) {
    const int64_t i = rows[k];
    const double xi = q_xyz[3 * i], yi = q_xyz[3 * i + 1], zi = q_xyz[3 * i + 2];
    int64_t w = out_ptr[k];
    for (int64_t j = full ? 0 : i + 1; j < pn; j++) {
      if (j == i) continue;
      const double dx = xi - q_xyz[3 * j], dy = yi - q_xyz[3 * j + 1], dz = zi - q_xyz[3 * j + 2];
      if (r2_search(dx, dy, dz) < sl2) out_list[w++] = (int32_t)j;
    }
  }
  return total;
}

/* ------------------------------------------------------------------------------------
 * Same contract, O(N): bin atoms into cubic cells of edge >= search length, scan the
 * 27-cell stencil, collect, sort each row ascending in j.
 * ---------------------------------------------------------------------------------- */
static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

int64_t ljo_makepair_cell(const double *q_xyz, int64_t pn, double search_len, int full,
                          int32_t *number_of_partners, int64_t *pointer,
                          int32_t *sorted_list, int64_t cap) {
  if (pn <= 0) return 0;
  const double sl2 = search_len * search_len;
  double lo[3] = {q_xyz[0], q_xyz[1], q_xyz[2]}, hi[3] = {q_xyz[0], q_xyz[1], q_xyz[2]};
  for (int64_t i = 1; i < pn; i++)
    for (int d = 0; d < 3; d++) {
      double v = q_xyz[3 * i + d];
      if (v < lo[d]) lo[d] = v;
      if (v > hi[d]) hi[d] = v;
    }
  int nc[3];
  for (int d = 0; d < 3; d++) {
    nc[d] = (int)floor((hi[d] - lo[d]) / search_len) + 1;
    if (nc[d] < 1) nc[d] = 1;
  }
  const int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
  int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)pn);
  int64_t *cstart = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
  int32_t *members = (int32_t *)malloc(sizeof(int32_t) * (size_t)pn);
  for (int64_t i = 0; i < pn; i++) {
    int c[3];
    for (int d = 0; d < 3; d++) {
      c[d] = (int)floor((q_xyz[3 * i + d] - lo[d]) / search_len);
      if (c[d] >= nc[d]) c[d] = nc[d] - 1;
      if (c[d] < 0) c[d] = 0;
    }
    cell_of[i] = (int32_t)(((int64_t)c[2] * nc[1] + c[1]) * nc[0] + c[0]);
    cstart[cell_of[i] + 1]++;
  }
  for (int64_t c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
  int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
  memcpy(cursor, cstart, sizeof(int64_t) * (size_t)ncell);
  for (int64_t i = 0; i < pn; i++) members[cursor[cell_of[i]]++] = (int32_t)i;
  free(cursor);

  /* pass 1: counts (parallel), pass 2: scan, pass 3: fill (parallel) */
  for (int pass = 0; pass < 2; pass++) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < pn; i++) {
      const double xi = q_xyz[3 * i], yi = q_xyz[3 * i + 1], zi = q_xyz[3 * i + 2];
      const int64_t ci = cell_of[i];
      const int cx = (int)(ci % nc[0]), cy = (int)((ci / nc[0]) % nc[1]),
                cz = (int)(ci / ((int64_t)nc[0] * nc[1]));
      int32_t cnt = 0;
      const int64_t base = pass ? pointer[i] : 0;
      const int fits = pass && (base + number_of_partners[i] <= cap);
      for (int dz = -1; dz <= 1; dz++) {
        const int z = cz + dz;
        if (z < 0 || z >= nc[2]) continue;
        for (int dy = -1; dy <= 1; dy++) {
          const int y = cy + dy;
          if (y < 0 || y >= nc[1]) continue;
          for (int dx = -1; dx <= 1; dx++) {
            const int x = cx + dx;
            if (x < 0 || x >= nc[0]) continue;
            const int64_t c = ((int64_t)z * nc[1] + y) * nc[0] + x;
            for (int64_t m = cstart[c]; m < cstart[c + 1]; m++) {
              const int32_t j = members[m];
              if (j == i || (!full && j < i)) continue;
              const double ddx = xi - q_xyz[3 * (int64_t)j], ddy = yi - q_xyz[3 * (int64_t)j + 1],
                           ddz = zi - q_xyz[3 * (int64_t)j + 2];
              if (r2_search(ddx, ddy, ddz) < sl2) {
                if (fits) sorted_list[base + cnt] = j;
                cnt++;
              }
            }
          }
        }
      }
      if (!pass) number_of_partners[i] = cnt;
      else if (fits) qsort(sorted_list + base, (size_t)cnt, sizeof(int32_t), cmp_i32);
    }
    if (!pass) {
      int64_t total = 0;
      for (int64_t i = 0; i < pn; i++) {
        pointer[i] = total;
        total += number_of_partners[i];
      }
    }
  }
  free(cell_of);
  free(cstart);
  free(members);
  int64_t total = pointer[pn - 1] + number_of_partners[pn - 1];
  return total <= cap ? total : -total;
}

/* ------------------------------------------------------------------------------------
 * Force loops.  q and p are addressed as base[c*comp_stride + i*elem_stride], which covers
 * AoS3 (comp 1, elem 3), AoS4 (comp 1, elem 4) and SoA planes (comp = plane stride,
 * elem 1; cpu_ref uses plane stride 400000, cpu_ref/force_soa.cpp:17-18).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t comp, elem;
} ljo_stride;

#define QX(i) q[(i)*qs.elem]
#define QY(i) q[qs.comp + (i)*qs.elem]
#define QZ(i) q[2 * qs.comp + (i)*qs.elem]
#define PX(i) p[(i)*ps.elem]
#define PY(i) p[ps.comp + (i)*ps.elem]
#define PZ(i) p[2 * ps.comp + (i)*ps.elem]

/* half list, i-major, register accumulation for i, reaction on j; pairs beyond the
 * cutoff are skipped (r2 > cl2 -> continue), cpu_ref/force_soa.cpp:163-195 */
void ljo_force_sorted(const double *q, int64_t q_comp, int64_t q_elem, double *p, int64_t p_comp,
                      int64_t p_elem, int64_t pn, double dt, double cl2,
                      const int32_t *sorted_list, const int32_t *number_of_partners,
                      const int64_t *pointer, int steps) {
  const ljo_stride qs = {q_comp, q_elem}, ps = {p_comp, p_elem};
  for (int s = 0; s < steps; s++) {
    for (int64_t i = 0; i < pn; i++) {
      const double xi = QX(i), yi = QY(i), zi = QZ(i);
      double fx = 0.0, fy = 0.0, fz = 0.0;
      const int64_t kp = pointer[i];
      const int32_t np = number_of_partners[i];
      for (int32_t k = 0; k < np; k++) {
        const int64_t j = sorted_list[kp + k];
        const double dx = QX(j) - xi, dy = QY(j) - yi, dz = QZ(j) - zi;
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > cl2) continue;
        const double r6 = r2 * r2 * r2;
        const double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2)) * dt;
        fx += df * dx;
        fy += df * dy;
        fz += df * dz;
        PX(j) -= df * dx;
        PY(j) -= df * dy;
        PZ(j) -= df * dz;
      }
      PX(i) += fx;
      PY(i) += fy;
      PZ(i) += fz;
    }
  }
}

/* full list gather: p[i] += sum_k df*d with df masked to 0 beyond the cutoff; rows are
 * independent, so this one is threaded (the all-cores CPU baseline). cuda/kernel.cuh:36-65 */
void ljo_force_gather(const double *q, int64_t q_comp, int64_t q_elem, double *p, int64_t p_comp,
                      int64_t p_elem, int64_t pn, double dt, double cl2,
                      const int32_t *sorted_list, const int32_t *number_of_partners,
                      const int64_t *pointer, int steps) {
  const ljo_stride qs = {q_comp, q_elem}, ps = {p_comp, p_elem};
  for (int s = 0; s < steps; s++) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < pn; i++) {
      const double xi = QX(i), yi = QY(i), zi = QZ(i);
      double fx = 0.0, fy = 0.0, fz = 0.0;
      const int64_t kp = pointer[i];
      const int32_t np = number_of_partners[i];
      for (int32_t k = 0; k < np; k++) {
        const int64_t j = sorted_list[kp + k];
        const double dx = QX(j) - xi, dy = QY(j) - yi, dz = QZ(j) - zi;
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double r6 = r2 * r2 * r2;
        double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2)) * dt;
        if (r2 > cl2) df = 0.0;
        fx += df * dx;
        fy += df * dy;
        fz += df * dz;
      }
      PX(i) += fx;
      PY(i) += fy;
      PZ(i) += fz;
    }
  }
}

/* The same `steps` applications for callers that KNOW q does not change between them (the
 * reference's benchmark: "static positions", cuda/force_cuda.cu:333-335 launches the kernel LOOP
 * times on the same q).  The per-step increment of row i is then the same double every step, so it
 * is evaluated once and accumulated `steps` times: bit-identical to ljo_force_gather (checked in
 * tests/test_oracle_cpu.py), at 1/steps of the cost -- this is what makes the full 100-step
 * comparison affordable at N = 1M on a small host. */
void ljo_force_gather_static(const double *q, int64_t q_comp, int64_t q_elem, double *p,
                             int64_t p_comp, int64_t p_elem, int64_t pn, double dt, double cl2,
                             const int32_t *sorted_list, const int32_t *number_of_partners,
                             const int64_t *pointer, int steps) {
  const ljo_stride qs = {q_comp, q_elem}, ps = {p_comp, p_elem};
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < pn; i++) {
    const double xi = QX(i), yi = QY(i), zi = QZ(i);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    const int64_t kp = pointer[i];
    const int32_t np = number_of_partners[i];
    for (int32_t k = 0; k < np; k++) {
      const int64_t j = sorted_list[kp + k];
      const double dx = QX(j) - xi, dy = QY(j) - yi, dz = QZ(j) - zi;
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double r6 = r2 * r2 * r2;
      double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2)) * dt;
      if (r2 > cl2) df = 0.0;
      fx += df * dx;
      fy += df * dy;
      fz += df * dz;
    }
    for (int s = 0; s < steps; s++) {
      PX(i) += fx;
      PY(i) += fy;
      PZ(i) += fz;
    }
  }
}

/* Gather for a SAMPLE of rows (see ljo_rows_brute): row k of the sample belongs to particle
 * rows[k], its entries are list[ptr[k] .. ptr[k] + nop[k]).  out_p[3k..3k+2] += the momentum of
 * that particle after `steps` applications (same arithmetic and order as ljo_force_gather). */
void ljo_force_rows(const double *q_xyz, const int64_t *rows, int64_t nrows, const int32_t *nop,
                    const int64_t *ptr, const int32_t *list, double dt, double cl2, int steps,
                    double *out_p) {
  for (int64_t k = 0; k < nrows; k++) {
    const int64_t i = rows[k];
    const double xi = q_xyz[3 * i], yi = q_xyz[3 * i + 1], zi = q_xyz[3 * i + 2];
    for (int s = 0; s < steps; s++) {
      double fx = 0.0, fy = 0.0, fz = 0.0;
      for (int32_t e = 0; e < nop[k]; e++) {
        const int64_t j = list[ptr[k] + e];
        const double dx = q_xyz[3 * j] - xi, dy = q_xyz[3 * j + 1] - yi, dz = q_xyz[3 * j + 2] - zi;
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double r6 = r2 * r2 * r2;
        double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2)) * dt;
        if (r2 > cl2) df = 0.0;
        fx += df * dx;
        fy += df * dy;
        fz += df * dz;
      }
      out_p[3 * k] += fx;
      out_p[3 * k + 1] += fy;
      out_p[3 * k + 2] += fz;
    }
  }
}

/* Column-major ELL gather (the reference's transposed_list, zero padded): entry k of row i
 * sits at list[i + k*pn]; cuda/kernel.cuh:102-134 */
void ljo_force_gather_ell(const double *q, int64_t q_comp, int64_t q_elem, double *p,
                          int64_t p_comp, int64_t p_elem, int64_t pn, double dt, double cl2,
                          const int32_t *transposed_list, const int32_t *number_of_partners,
                          int steps) {
  const ljo_stride qs = {q_comp, q_elem}, ps = {p_comp, p_elem};
  for (int s = 0; s < steps; s++) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < pn; i++) {
      const double xi = QX(i), yi = QY(i), zi = QZ(i);
      double fx = 0.0, fy = 0.0, fz = 0.0;
      const int32_t np = number_of_partners[i];
      for (int32_t k = 0; k < np; k++) {
        const int64_t j = transposed_list[i + (int64_t)k * pn];
        const double dx = QX(j) - xi, dy = QY(j) - yi, dz = QZ(j) - zi;
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double r6 = r2 * r2 * r2;
        double df = ((24.0 * r6 - 48.0) / (r6 * r6 * r2)) * dt;
        if (r2 > cl2) df = 0.0;
        fx += df * dx;
        fy += df * dy;
        fz += df * dz;
      }
      PX(i) += fx;
      PY(i) += fy;
      PZ(i) += fz;
    }
  }
}

/* Per-row shuffle with one mt19937(seed) stream, rows visited in order, each row shuffled
 * the way libstdc++'s std::shuffle does for a 32-bit engine (pairs of swaps drawn from one
 * 32-bit word while the product of the two ranges fits).  cuda/force_cuda.cu:255-263 */
void ljo_shuffle_rows(int32_t *sorted_list, const int32_t *number_of_partners,
                      const int64_t *pointer, int64_t pn, uint32_t seed) {
  ljo_mt g;
  mt_seed(&g, seed);
  for (int64_t r = 0; r < pn; r++) {
    int32_t *a = sorted_list + pointer[r];
    const uint32_t n = (uint32_t)number_of_partners[r];
    if (n < 2) continue;
    /* urange = n-1 <= urngrange/urange always holds for the row lengths seen here
       (n*n < 2^32), so libstdc++ takes the two-at-a-time path. */
    uint32_t i = 1;
    if ((n % 2u) == 0u) {
      uint32_t pos = mt_below(&g, 2u);
      int32_t t = a[1];
      a[1] = a[pos];
      a[pos] = t;
      i = 2;
    }
    while (i < n) {
      const uint32_t swap_range = i + 1u;
      const uint32_t b1 = swap_range + 1u;
      const uint32_t x = mt_below(&g, swap_range * b1);
      const uint32_t p0 = x / b1, p1 = x % b1;
      int32_t t = a[i];
      a[i] = a[p0];
      a[p0] = t;
      t = a[i + 1];
      a[i + 1] = a[p1];
      a[p1] = t;
      i += 2;
    }
  }
}

/* CSR -> column-major ELL, padding value 0.  Returns max row length.  cuda/force_cuda.cu:229-240 */
int32_t ljo_transpose_list(const int32_t *sorted_list, const int32_t *number_of_partners,
                           const int64_t *pointer, int64_t pn, int32_t *transposed_list,
                           int64_t cap_entries) {
  int32_t max_np = 0;
  for (int64_t i = 0; i < pn; i++)
    if (number_of_partners[i] > max_np) max_np = number_of_partners[i];
  if ((int64_t)max_np * pn > cap_entries) return -max_np;
  memset(transposed_list, 0, sizeof(int32_t) * (size_t)((int64_t)max_np * pn));
  for (int64_t i = 0; i < pn; i++)
    for (int32_t k = 0; k < number_of_partners[i]; k++)
      transposed_list[i + (int64_t)k * pn] = sorted_list[pointer[i] + k];
  return max_np;
}

int ljo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ljo_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
