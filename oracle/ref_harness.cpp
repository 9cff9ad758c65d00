// ref_harness.cpp -- thin C-ABI window onto the UNMODIFIED reference cpu_ref/force_soa.cpp.
// TEST INFRASTRUCTURE ONLY (see oracle/lj_oracle.c header for who may load it).
//
// The reference translation unit is compiled straight from /root/reference by
// oracle/Makefile (its `main` renamed with -Dmain=lj_ref_main, and `density` selected by
// a sed on the compiler's stdin -- no reference source is ever written to disk here).
// All of its state has external linkage (cpu_ref/force_soa.cpp:10-28), so this file only
// declares it and hands out pointers.  It also exposes the real libstdc++ std::shuffle,
// which cuda/force_cuda.cu:255-263 uses for random_shfl(), so that the C restatement in
// lj_oracle.c can be pinned against it.
#include <algorithm>
#include <cstdint>
#include <random>

// ---- the reference's globals and functions (cpu_ref/force_soa.cpp) ----
extern double q[4][400000];
extern double p[4][400000];
extern double L;
extern int particle_number;
extern int number_of_pairs;
extern int number_of_partners[];
extern int i_particles[];
extern int j_particles[];
extern int pointer[];
extern int sorted_list[];
void init(void);
void makepair(void);
void sortpair(void);
void force_pair(void);
void force_sorted(void);
void force_next(void);
void force_intrin(void);
void savepair(void);
void loadpair(void);

extern "C" {

double ljref_density(void) { return LJREF_DENSITY; }
void ljref_set_L(double l) { L = l; }
double ljref_get_L(void) { return L; }
int ljref_plane_stride(void) { return 400000; }
// init() owns a function-static mt19937: call it once per process.
void ljref_init(void) { init(); }
void ljref_makepair(void) {
  number_of_pairs = 0;
  makepair();
}
void ljref_sortpair(void) { sortpair(); }
// the reference's own pair.dat writer / reader (cpu_ref/force_soa.cpp:360-377), in the CWD
void ljref_savepair(void) {
  number_of_pairs = 0;
  savepair();
}
void ljref_loadpair(void) { loadpair(); }
void ljref_zero_p(void) {
  for (int c = 0; c < 4; c++) std::fill(p[c], p[c] + 400000, 0.0);
}
// kind: 0 pair, 1 sorted, 2 next, 3 intrin -- `steps` back-to-back calls, like measure()
void ljref_force(int kind, int steps) {
  void (*f)(void) = kind == 0 ? force_pair : kind == 1 ? force_sorted : kind == 2 ? force_next : force_intrin;
  for (int s = 0; s < steps; s++) f();
}
int ljref_particle_number(void) { return particle_number; }
int ljref_number_of_pairs(void) { return number_of_pairs; }
double* ljref_q(void) { return &q[0][0]; }
double* ljref_p(void) { return &p[0][0]; }
int* ljref_number_of_partners(void) { return number_of_partners; }
int* ljref_pointer(void) { return pointer; }
int* ljref_sorted_list(void) { return sorted_list; }
int* ljref_i_particles(void) { return i_particles; }
int* ljref_j_particles(void) { return j_particles; }

// libstdc++ std::shuffle per row with one mt19937(seed) stream (what random_shfl() runs)
void ljref_std_shuffle_rows(int32_t* list, const int32_t* nop, const int64_t* ptr, int64_t pn,
                            uint32_t seed) {
  std::mt19937 mt(seed);
  for (int64_t i = 0; i < pn; i++) std::shuffle(list + ptr[i], list + ptr[i] + nop[i], mt);
}

}  // extern "C"
