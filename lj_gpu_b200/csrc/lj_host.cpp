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
