// lj_host.cpp -- host-side helpers of the C ABI that involve no device work.
//
// lj_init_fcc replaces init()/add_particle() of the reference driver
// (cuda/force_cuda.cu:47-94): the benchmark's input is generated in-process, so a drop-in
// must produce bit-identical positions.  That pins three things: the lattice arithmetic
// (s = (rho/4)^(-1/3), n = int(L/s) cells per side, FCC basis), the visiting order
// (z, then y, then x, then the four basis atoms) and the jitter source (ONE std::mt19937
// seeded with 2, three uniform_real_distribution<double>(0, 0.1) draws per atom: x, y, z).
#include <cmath>
#include <cstdint>
#include <random>

#include "../../include/lj_b200.h"

extern "C" int64_t lj_init_fcc(double density, double L, double* q_xyz_host, int64_t cap_particles,
                               int32_t* cells_per_side_out) {
  if (!(density > 0.0) || !(L > 0.0)) return 0;
  const double spacing = 1.0 / std::pow(density * 0.25, 1.0 / 3.0);
  const double half = spacing * 0.5;
  const int n = static_cast<int>(L / spacing);
  if (cells_per_side_out) *cells_per_side_out = n;
  const int64_t need = 4LL * n * n * n;
  if (need > cap_particles || !q_xyz_host) return -need;

  std::mt19937 engine(2);
  std::uniform_real_distribution<double> jitter(0.0, 0.1);
  // basis atom b sits at (ox,oy,oz) half-spacings from the cell corner
  static const int basis[4][3] = {{0, 0, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
  double* out = q_xyz_host;
  for (int iz = 0; iz < n; ++iz) {
    const double z0 = iz * spacing;
    for (int iy = 0; iy < n; ++iy) {
      const double y0 = iy * spacing;
      for (int ix = 0; ix < n; ++ix) {
        const double x0 = ix * spacing;
        for (const auto& b : basis) {
          const double bx = b[0] ? x0 + half : x0;
          const double by = b[1] ? y0 + half : y0;
          const double bz = b[2] ? z0 + half : z0;
          out[0] = bx + jitter(engine);
          out[1] = by + jitter(engine);
          out[2] = bz + jitter(engine);
          out += 3;
        }
      }
    }
  }
  return need;
}

// ------------------------------------------------------------------------------------------
// Pair-list cache files of the reference, so lists can be exchanged with the
// reference binaries.  Host-only, no device work.
//
//  * text cache `.cache_pair_{all,half}.dat` of cuda/force_cuda.cu (makepaircache :165-176,
//    loadpair :203-227): "pn npairs\n", then pn lines "number_of_partners pointer", then npairs
//    lines "j".  The reader applies check_loadedpair()'s range checks (:183-201).
//  * binary `pair.dat` of cpu_ref (savepair/loadpair, cpu_ref/force_soa.cpp:360-377):
//    int npairs; int number_of_partners[N]; int i_particles[MAX_PAIRS]; int j_particles[MAX_PAIRS]
//    with the reference's compile-time N = 400000 and MAX_PAIRS = 30*N (97.6 MB, mostly padding).
// ------------------------------------------------------------------------------------------
#include <cstdio>
#include <vector>

extern "C" int lj_paircache_write_text(const char* path, int64_t pn, int64_t npairs,
                                       const int32_t* number_of_partners, const int32_t* pointer,
                                       const int32_t* sorted_list) {
  if (!path || pn < 0 || npairs < 0 || (pn && (!number_of_partners || !pointer)) || (npairs && !sorted_list))
    return LJ_ERR_BAD_ARG;
  FILE* fp = std::fopen(path, "w");
  if (!fp) return LJ_ERR_BAD_ARG;
  std::fprintf(fp, "%d %d\n", (int)pn, (int)npairs);
  for (int64_t i = 0; i < pn; i++) std::fprintf(fp, "%d %d\n", number_of_partners[i], pointer[i]);
  for (int64_t k = 0; k < npairs; k++) std::fprintf(fp, "%d\n", sorted_list[k]);
  const bool ok = std::fclose(fp) == 0;
  return ok ? LJ_OK : LJ_ERR_BAD_ARG;
}

// pn_expected < 0: accept any particle count.  Arrays may be NULL to query the sizes only.
extern "C" int lj_paircache_read_text(const char* path, int64_t pn_expected, int64_t* pn_out,
                                      int64_t* npairs_out, int32_t* number_of_partners,
                                      int32_t* pointer, int64_t cap_particles, int32_t* sorted_list,
                                      int64_t cap_pairs) {
  if (!path || !pn_out || !npairs_out) return LJ_ERR_BAD_ARG;
  FILE* fp = std::fopen(path, "r");
  if (!fp) return LJ_ERR_BAD_ARG;
  int pn = 0, npairs = 0;
  int rc = LJ_OK;
  if (std::fscanf(fp, "%d %d", &pn, &npairs) != 2 || pn < 0 || npairs < 0) rc = LJ_ERR_INVALID_LIST;
  *pn_out = pn; *npairs_out = npairs;
  if (rc == LJ_OK && pn_expected >= 0 && pn != pn_expected) rc = LJ_ERR_INVALID_LIST;  // "may be broken"
  if (rc == LJ_OK && number_of_partners && pointer && sorted_list) {
    if (pn > cap_particles || npairs > cap_pairs) rc = LJ_ERR_CAPACITY;
    for (int i = 0; rc == LJ_OK && i < pn; i++) {
      int n, p;
      if (std::fscanf(fp, "%d %d", &n, &p) != 2 || n < 0 || n >= pn || p < 0 || p > npairs) rc = LJ_ERR_INVALID_LIST;
      else { number_of_partners[i] = n; pointer[i] = p; }
    }
    for (int k = 0; rc == LJ_OK && k < npairs; k++) {
      int j;
      if (std::fscanf(fp, "%d", &j) != 1 || j < 0 || j >= pn) rc = LJ_ERR_INVALID_LIST;
      else sorted_list[k] = j;
    }
  }
  std::fclose(fp);
  return rc;
}

extern "C" int lj_pairdat_write(const char* path, int64_t n_static, int64_t max_pairs_static,
                                int64_t pn, int64_t npairs, const int32_t* number_of_partners,
                                const int32_t* i_particles, const int32_t* j_particles) {
  if (!path || pn > n_static || npairs > max_pairs_static || pn < 0 || npairs < 0) return LJ_ERR_BAD_ARG;
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return LJ_ERR_BAD_ARG;
  const int32_t np32 = (int32_t)npairs;
  std::vector<int32_t> pad((size_t)(n_static > max_pairs_static ? n_static : max_pairs_static), 0);
  bool ok = std::fwrite(&np32, 4, 1, fp) == 1;
  auto put = [&](const int32_t* src, int64_t n, int64_t total) {
    ok = ok && (n == 0 || std::fwrite(src, 4, (size_t)n, fp) == (size_t)n);
    ok = ok && (total == n || std::fwrite(pad.data(), 4, (size_t)(total - n), fp) == (size_t)(total - n));
  };
  put(number_of_partners, pn, n_static);
  put(i_particles, npairs, max_pairs_static);
  put(j_particles, npairs, max_pairs_static);
  ok = (std::fclose(fp) == 0) && ok;
  return ok ? LJ_OK : LJ_ERR_BAD_ARG;
}

extern "C" int lj_pairdat_read(const char* path, int64_t n_static, int64_t max_pairs_static, int64_t pn,
                               int64_t* npairs_out, int32_t* number_of_partners, int32_t* i_particles,
                               int32_t* j_particles, int64_t cap_pairs) {
  if (!path || !npairs_out || pn < 0 || pn > n_static) return LJ_ERR_BAD_ARG;
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return LJ_ERR_BAD_ARG;
  int32_t np32 = 0;
  int rc = std::fread(&np32, 4, 1, fp) == 1 && np32 >= 0 && np32 <= max_pairs_static ? LJ_OK : LJ_ERR_INVALID_LIST;
  *npairs_out = np32;
  if (rc == LJ_OK && number_of_partners && i_particles && j_particles) {
    if (np32 > cap_pairs) rc = LJ_ERR_CAPACITY;
    auto get = [&](int32_t* dst, int64_t n, int64_t index_of_block) {
      if (rc != LJ_OK) return;
      if (std::fseek(fp, (long)(4 + 4 * index_of_block), SEEK_SET) != 0 ||
          (n && std::fread(dst, 4, (size_t)n, fp) != (size_t)n))
        rc = LJ_ERR_INVALID_LIST;
    };
    get(number_of_partners, pn, 0);
    get(i_particles, np32, n_static);
    get(j_particles, np32, n_static + max_pairs_static);
    for (int64_t k = 0; rc == LJ_OK && k < np32; k++)
      if (i_particles[k] < 0 || i_particles[k] >= pn || j_particles[k] < 0 || j_particles[k] >= pn)
        rc = LJ_ERR_INVALID_LIST;
  }
  std::fclose(fp);
  return rc;
}
