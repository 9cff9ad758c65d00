// force_b200.cpp -- C++ host driver: the reference's benchmark program on the new library.
//
// Keeps the observable behaviour of cuda/force_cuda.cu main()/measure() (:319-444):
//   * same input (jittered FCC lattice from one mt19937(2) stream), same constants
//     (density 0.5, L 50, dt 0.001, cutoff 3.0, search 3.3, LOOP 100),
//   * `./force_b200 [THREAD_BLOCK]` stays valid (first positional argument, 64..1024),
//   * stderr carries the two timing lines of measure() verbatim,
//     "N=%d, %s %f [sec]" and "N=%d, %s %f [sec] (without Host<->Device)",
//   * with --test (the EN_TEST_GPU build) stdout is print_results(): p[0..4], p[pn-5..pn-1]
//     in "%.10f %.10f %.10f", comparable with ref_data/density*.dat.
// What changed underneath: makepair() is lj_build_list on the GPU (no pair cache needed),
// the kernels are the sm_100a ones behind lj_force_step, memory comes from the library.
//
// Extra flags (the reference has compile-time constants instead, SURVEY 5 "Config"):
//   --density R  --L X  --layout aos3|aos4|soa  --variant auto|warp|thread|subwarp|tile|n3
//   --group G  --prec fp64|mixed  --steps K  --rebuild-every M  --graph  --test  --all
//   --cache   use / create the reference's text pair cache .cache_pair_{all,half}.dat in the CWD
//   --soa6    the OpenACC SoA program (openacc/force_oacc_soa.cpp): six separate arrays, CSR run
//             (acc_reactless_soa) then transposed-list run (acc_reactless_memopt_soa), each followed
//             by print_results()
//             (loadpair()/makepaircache(), cuda/force_cuda.cu:165-227): a cached list is uploaded
//             like the reference does, otherwise the list is built on the GPU and written out
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lj_b200.h"

namespace {

struct Options {
  int thread_block = 128;
  double density = 0.5, L = 50.0;
  std::string layout = "aos4", variant = "auto", prec = "fp64";
  int group = 0, steps = 100, rebuild_every = 0;
  bool graph = false, test = false, all = false, cache = false, soa6 = false, print = false, md = false;
  int gpus = 1;
  bool one_device = false;  // --one-device: every slab of --gpus N on device 0 (the decomposed path on a one-GPU box)
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

[[noreturn]] void die(lj_ctx* ctx, int rc, const char* where) {
  std::fprintf(stderr, "%s: %s: %s\n", where, lj_status_string(rc), ctx ? lj_last_error_string(ctx) : "");
  std::exit(EXIT_FAILURE);  // the driver, not the library, decides to exit
}

int layout_id(const std::string& s) {
  if (s == "aos3") return LJ_AOS_D3;
  if (s == "aos4") return LJ_AOS_D4;
  if (s == "soa") return LJ_SOA_D;
  std::fprintf(stderr, "unknown layout %s\n", s.c_str());
  std::exit(1);
}

// q (packed xyz) -> the requested layout; p zero-initialised like init() does
void pack(const std::vector<double>& xyz, int64_t pn, int layout, std::vector<double>& q,
          std::vector<double>& p) {
  const int w = layout == LJ_AOS_D4 ? 4 : 3;
  q.assign((size_t)pn * w, 0.0);
  p.assign((size_t)pn * w, 0.0);
  for (int64_t i = 0; i < pn; i++)
    for (int c = 0; c < 3; c++) {
      if (layout == LJ_SOA_D) q[(size_t)c * pn + i] = xyz[3 * i + c];
      else q[(size_t)i * w + c] = xyz[3 * i + c];
    }
}

void print_results(const std::vector<double>& p, int64_t pn, int layout) {
  const int w = layout == LJ_AOS_D4 ? 4 : 3;
  auto at = [&](int64_t i, int c) { return layout == LJ_SOA_D ? p[(size_t)c * pn + i] : p[(size_t)i * w + c]; };
  for (int64_t i = 0; i < 5; i++) std::fprintf(stdout, "%.10f %.10f %.10f\n", at(i, 0), at(i, 1), at(i, 2));
  for (int64_t i = pn - 5; i < pn; i++) std::fprintf(stdout, "%.10f %.10f %.10f\n", at(i, 0), at(i, 1), at(i, 2));
}

// build the list once more on the GPU into caller-owned arrays and write the reference's cache
void write_cache(lj_ctx* ctx, const char* file, const std::vector<double>& xyz, int64_t pn, int half) {
  void *q = nullptr, *nop = nullptr, *ptr = nullptr, *list = nullptr;
  lj_list_args l{};
  int rc = lj_dev_alloc(ctx, (size_t)pn * 24, &q, nullptr);
  if (!rc) rc = lj_dev_alloc(ctx, (size_t)pn * 4, &nop, nullptr);
  if (!rc) rc = lj_dev_alloc(ctx, (size_t)pn * 4, &ptr, nullptr);
  if (!rc) rc = lj_upload(ctx, q, xyz.data(), (size_t)pn * 24, nullptr);
  l.q = q; l.pn = pn; l.layout = LJ_AOS_D3; l.half = half; l.search_len = 3.3;
  l.number_of_partners = (int32_t*)nop; l.pointer = ptr; l.flags = LJ_LIST_SORT_ROWS;
  int64_t npairs = 0;
  if (!rc) { rc = lj_build_list(ctx, &l, &npairs, nullptr); if (rc == LJ_ERR_CAPACITY) rc = LJ_OK; }
  if (!rc) rc = lj_dev_alloc(ctx, (size_t)(npairs ? npairs : 1) * 4, &list, nullptr);
  l.sorted_list = (int32_t*)list; l.capacity = npairs;
  if (!rc) rc = lj_build_list(ctx, &l, &npairs, nullptr);
  std::vector<int32_t> h_nop(pn), h_ptr(pn), h_list(npairs ? npairs : 1);
  if (!rc) rc = lj_download(ctx, h_nop.data(), nop, (size_t)pn * 4, nullptr);
  if (!rc) rc = lj_download(ctx, h_ptr.data(), ptr, (size_t)pn * 4, nullptr);
  if (!rc) rc = lj_download(ctx, h_list.data(), list, (size_t)npairs * 4, nullptr);
  if (!rc) rc = lj_sync(ctx, nullptr);
  if (!rc) rc = lj_paircache_write_text(file, pn, npairs, h_nop.data(), h_ptr.data(), h_list.data());
  lj_dev_free(ctx, q, nullptr); lj_dev_free(ctx, nop, nullptr); lj_dev_free(ctx, ptr, nullptr);
  lj_dev_free(ctx, list, nullptr);
  if (rc) die(ctx, rc, "write_cache");
}

void measure(lj_ctx* ctx, const Options& o, const std::vector<double>& xyz, int64_t pn,
             const std::string& layout, const std::string& variant, int group, const char* name,
             bool print) {
  const int lay = layout_id(layout);
  std::vector<double> q, p;
  pack(xyz, pn, lay, q, p);
  lj_measure_args m{};
  m.q_host = q.data(); m.p_host = p.data(); m.pn = pn; m.layout = lay;
  m.plane_stride = lay == LJ_SOA_D ? pn : 0;
  m.dt = 0.001; m.cl2 = 3.0 * 3.0; m.search_len = 3.3;
  m.loop = o.steps; m.rebuild_every = o.rebuild_every;
  m.half = variant == "n3";
  m.variant = variant == "tile" ? LJ_VARIANT_TILE_TMA : variant == "n3" ? LJ_VARIANT_NEWTON3
              : variant == "auto" ? LJ_VARIANT_AUTO : LJ_VARIANT_SUBWARP;
  m.group = group ? group : (variant == "warp" ? 32 : variant == "thread" ? 1 : 0);
  m.precision = o.prec == "mixed" ? LJ_PREC_MIXED : LJ_PREC_FP64;
  m.threads_per_block = o.thread_block; m.use_graph = o.graph;
  // pair cache of the reference driver: same file names, same text format
  const char* cache_file = m.half ? ".cache_pair_half.dat" : ".cache_pair_all.dat";
  std::vector<int32_t> c_nop, c_ptr, c_list;
  if (o.cache) {
    int64_t cpn = 0, cnp = 0;
    if (lj_paircache_read_text(cache_file, pn, &cpn, &cnp, nullptr, nullptr, 0, nullptr, 0) == LJ_OK) {
      c_nop.resize(cpn); c_ptr.resize(cpn); c_list.resize(cnp ? cnp : 1);
      if (lj_paircache_read_text(cache_file, pn, &cpn, &cnp, c_nop.data(), c_ptr.data(), cpn, c_list.data(),
                                 cnp) == LJ_OK) {
        std::fprintf(stderr, "%s is successfully loaded.\n", cache_file);
        m.list_host = c_list.data(); m.number_of_partners_host = c_nop.data(); m.pointer_host = c_ptr.data();
        m.number_of_pairs_in = cnp;
      } else {
        std::fprintf(stderr, "Pairlist cache data may be broken.\n");
      }
    } else {
      std::fprintf(stderr, "Now make pairlist %s.\n", cache_file);
    }
  }
  const int rc = lj_measure(ctx, &m);
  if (rc) die(ctx, rc, "lj_measure");
  if (o.cache && !m.list_host) write_cache(ctx, cache_file, xyz, pn, m.half);
  std::fprintf(stderr, "N=%d, %s %f [sec]\n", (int)pn, name, m.seconds_total);
  std::fprintf(stderr, "N=%d, %s %f [sec] (without Host<->Device)\n", (int)pn, name, m.seconds_kernel);
  std::fprintf(stderr, "  pairs=%lld max_partners=%d list_builds=%d  %.4g pair-interactions/s\n",
               (long long)m.number_of_pairs, m.max_partners, m.list_builds,
               (double)m.number_of_pairs * (m.half ? 2.0 : 1.0) * o.steps / m.seconds_kernel);
  if (lj_list_mirror_token(ctx))
    std::fprintf(stderr, "  cell-tile mirror of the %s list: yes\n", m.list_host ? "loaded" : "built");
  if (print) print_results(p, pn, lay);
}

// --gpus N: the same benchmark on N GPUs of this box from ONE process (lj_decomp_*): z-slabs of lattice
// layers, ghost positions pulled over NVLink every step, interior tiles overlapped with the pull.
// No counterpart in the reference (single GPU); timing lines in the reference's format.
int run_decomposed(const Options& o, const std::vector<double>& xyz, int64_t pn) {
  std::vector<int64_t> slab((size_t)o.gpus + 1);
  int64_t halo_rows = 0;
  int rc = lj_decomp_plan_fcc(o.density, o.L, o.gpus, 3.3, slab.data(), &halo_rows);
  if (rc) {
    std::fprintf(stderr, "lj_decomp_plan_fcc: %s (a slab thinner than the halo: fewer GPUs or a larger box)\n", lj_status_string(rc));
    return 1;
  }
  lj_decomp_args a{};
  a.ngpus = o.gpus; a.q_xyz_host = xyz.data(); a.pn = pn; a.slab_begin = slab.data(); a.halo_rows = halo_rows;
  a.search_len = 3.3; a.cutoff = 3.0; a.dt = 0.001;
  a.precision = o.prec == "mixed" ? LJ_PREC_MIXED : LJ_PREC_FP64;
  const std::vector<int32_t> dev0((size_t)o.gpus, 0);
  if (o.one_device) a.devices = dev0.data();
  lj_decomp* d = nullptr;
  const double t_all = now();
  rc = lj_decomp_create(&d, &a);
  if (rc) { std::fprintf(stderr, "lj_decomp_create: %s: %s\n", lj_status_string(rc), lj_decomp_last_error(d)); return 1; }
  const double t_calc = now();
  rc = o.md ? lj_decomp_md(d, o.steps, o.rebuild_every, 1) : lj_decomp_step(d, o.steps, o.rebuild_every, 1);
  if (!rc) rc = lj_decomp_sync(d);
  const double sec = now() - t_calc;
  std::vector<double> p((size_t)pn * 3);
  if (!rc) rc = lj_decomp_gather(d, p.data(), nullptr);
  if (rc) { std::fprintf(stderr, "lj_decomp: %s: %s\n", lj_status_string(rc), lj_decomp_last_error(d)); return 1; }
  const char* name = o.md ? "force_decomposed_md" : "force_decomposed";
  std::fprintf(stderr, "N=%d, %s_%dgpus %f [sec]\n", (int)pn, name, o.gpus, now() - t_all);
  std::fprintf(stderr, "N=%d, %s_%dgpus %f [sec] (without Host<->Device)\n", (int)pn, name, o.gpus, sec);
  std::fprintf(stderr, "  pairs=%lld slabs=%d halo_rows=%lld  %.4g pair-interactions/s  launches=%lld\n",
               (long long)lj_decomp_pairs(d), o.gpus, (long long)halo_rows,
               (double)lj_decomp_pairs(d) * o.steps / sec, (long long)lj_decomp_launch_count(d));
  if (o.print) print_results(p, pn, LJ_AOS_D3);
  lj_decomp_destroy(d);
  return 0;
}

// The OpenACC SoA program (openacc/force_oacc_soa.cpp) on the new library: six separately
// allocated arrays qx,qy,qz,px,py,pz (:17-22), makepair + CSR list (force_reactless, OACC_REF) or
// the transposed list (force_reactless_memopt, OACC_TRANS), LOOP force calls between one upload
// and one download (measure(), :265-296), print_results() (:298-306).
void measure_soa6(lj_ctx* ctx, const Options& o, const std::vector<double>& xyz, int64_t pn, bool transposed) {
  std::vector<double> h[3];
  for (int c = 0; c < 3; c++) {
    h[c].resize((size_t)pn);
    for (int64_t i = 0; i < pn; i++) h[c][(size_t)i] = xyz[3 * i + c];
  }
  void* q[3] = {nullptr, nullptr, nullptr};
  void* p[3] = {nullptr, nullptr, nullptr};
  void *nop = nullptr, *ptr = nullptr, *list = nullptr, *tl = nullptr, *spacer = nullptr;
  const size_t b = (size_t)pn * sizeof(double);
  int rc = LJ_OK;
  for (int c = 0; c < 3 && !rc; c++) {
    rc = lj_dev_alloc(ctx, b, &q[c], nullptr);
    if (!rc && c == 0) rc = lj_dev_alloc(ctx, 4096 + 8 * 17, &spacer, nullptr);  // no regular spacing
    if (!rc) rc = lj_dev_alloc(ctx, b, &p[c], nullptr);
    if (!rc) rc = lj_upload(ctx, q[c], h[c].data(), b, nullptr);
  }
  std::vector<double> zero((size_t)pn, 0.0);
  for (int c = 0; c < 3 && !rc; c++) rc = lj_upload(ctx, p[c], zero.data(), b, nullptr);
  if (!rc) rc = lj_dev_alloc(ctx, (size_t)pn * 4, &nop, nullptr);
  if (!rc) rc = lj_dev_alloc(ctx, (size_t)pn * 4, &ptr, nullptr);
  if (rc) die(ctx, rc, "measure_soa6: allocate");
  lj_list_args l{};
  l.pn = pn; l.search_len = 3.3;
  l.number_of_partners = (int32_t*)nop; l.pointer = ptr;
  int64_t npairs = 0;
  rc = lj_build_list_soa6(ctx, (double*)q[0], (double*)q[1], (double*)q[2], &l, &npairs, nullptr);
  if (rc == LJ_ERR_CAPACITY) rc = LJ_OK;  // first call sizes the list
  if (!rc) rc = lj_dev_alloc(ctx, (size_t)(npairs ? npairs : 1) * 4, &list, nullptr);
  l.sorted_list = (int32_t*)list; l.capacity = npairs;
  if (!rc) rc = lj_build_list_soa6(ctx, (double*)q[0], (double*)q[1], (double*)q[2], &l, &npairs, nullptr);
  if (rc) die(ctx, rc, "lj_build_list_soa6");
  lj_force_args a{};
  a.pn = pn; a.dt = 0.001; a.cl2 = 3.0 * 3.0;
  a.number_of_partners = (int32_t*)nop;
  a.variant = LJ_VARIANT_AUTO;
  a.precision = o.prec == "mixed" ? LJ_PREC_MIXED : LJ_PREC_FP64;
  a.threads_per_block = o.thread_block;
  if (transposed) {  // make_transposed_list() (:143-154)
    int32_t max_np = 0;
    rc = lj_list_result(ctx, nullptr, &max_np, nullptr);
    if (!rc) rc = lj_dev_alloc(ctx, (size_t)(max_np ? max_np : 1) * (size_t)pn * 4, &tl, nullptr);
    if (!rc) rc = lj_build_ell(ctx, (int32_t*)list, (int32_t*)nop, ptr, 0, pn, (int32_t*)tl, (int64_t)max_np * pn, &max_np, nullptr);
    if (rc) die(ctx, rc, "lj_build_ell");
    a.list = (int32_t*)tl; a.pointer = nullptr; a.list_layout = LJ_LIST_ELL;
  } else {
    a.list = (int32_t*)list; a.pointer = ptr; a.list_layout = LJ_LIST_CSR; a.list_entries = npairs;
  }
  lj_sync(ctx, nullptr);
  const auto t0 = std::chrono::steady_clock::now();
  rc = lj_force_loop_soa6(ctx, (double*)q[0], (double*)q[1], (double*)q[2], (double*)p[0], (double*)p[1],
                          (double*)p[2], &a, o.steps, nullptr);
  if (!rc) rc = lj_sync(ctx, nullptr);
  if (rc) die(ctx, rc, "lj_force_loop_soa6");
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::fprintf(stderr, "N=%d, %s %f [sec] (without Host<->Device)\n", (int)pn,
               transposed ? "acc_reactless_memopt_soa" : "acc_reactless_soa", sec);
  for (int c = 0; c < 3 && !rc; c++) rc = lj_download(ctx, h[c].data(), p[c], b, nullptr);
  if (!rc) rc = lj_sync(ctx, nullptr);
  if (rc) die(ctx, rc, "measure_soa6: download");
  for (int64_t i = 0; i < 5; i++) std::fprintf(stdout, "%.10f %.10f %.10f\n", h[0][i], h[1][i], h[2][i]);
  for (int64_t i = pn - 5; i < pn; i++) std::fprintf(stdout, "%.10f %.10f %.10f\n", h[0][i], h[1][i], h[2][i]);
  for (void* d : {q[0], q[1], q[2], p[0], p[1], p[2], nop, ptr, list, tl, spacer})
    if (d) lj_dev_free(ctx, d, nullptr);
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  for (int a = 1; a < argc; a++) {
    std::string s = argv[a];
    auto next = [&]() -> const char* {
      if (a + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", s.c_str()); std::exit(1); }
      return argv[++a];
    };
    if (s == "--density") o.density = std::atof(next());
    else if (s == "--L") o.L = std::atof(next());
    else if (s == "--layout") o.layout = next();
    else if (s == "--variant") o.variant = next();
    else if (s == "--group") o.group = std::atoi(next());
    else if (s == "--prec") o.prec = next();
    else if (s == "--steps") o.steps = std::atoi(next());
    else if (s == "--rebuild-every") o.rebuild_every = std::atoi(next());
    else if (s == "--graph") o.graph = true;
    else if (s == "--test") o.test = true;
    else if (s == "--all") o.all = true;
    else if (s == "--cache") o.cache = true;
    else if (s == "--print") o.print = true;
    else if (s == "--gpus") o.gpus = std::atoi(next());
    else if (s == "--md") o.md = true;
    else if (s == "--one-device") o.one_device = true;
    else if (s == "--soa6") o.soa6 = true;
    else if (s[0] != '-') o.thread_block = std::atoi(s.c_str());
    else { std::fprintf(stderr, "unknown option %s\n", s.c_str()); return 1; }
  }
  if (o.thread_block < 64 || o.thread_block > 1024) {  // cuda/force_cuda.cu:380-383
    std::fprintf(stderr, "THREAD_BLOCK size is too large or small.\n");
    return 1;
  }

  int cells = 0;
  const int64_t need = -lj_init_fcc(o.density, o.L, nullptr, 0, &cells);
  std::vector<double> xyz((size_t)need * 3);
  const int64_t pn = lj_init_fcc(o.density, o.L, xyz.data(), need, &cells);
  if (pn <= 0) { std::fprintf(stderr, "empty system\n"); return 1; }

  if (o.gpus > 1) return run_decomposed(o, xyz, pn);

  lj_ctx* ctx = nullptr;
  int rc = lj_ctx_create(&ctx, 0);
  if (rc) die(nullptr, rc, "lj_ctx_create");

  if (o.soa6) {
    // openacc/force_oacc_soa.cpp main(): OACC_REF (CSR list) then OACC_TRANS (transposed list);
    // each prints the ten momenta of ref_data/density*.dat
    measure_soa6(ctx, o, xyz, pn, false);
    measure_soa6(ctx, o, xyz, pn, true);
  } else if (o.test) {
    // the EN_TEST_GPU build: one kernel for double3 then double4, then print p_d3.  Unlike the
    // reference (which re-uploads the accumulated p, force_cuda.cu:331,338) each measure()
    // here starts from p = 0, so the printed values are those of ONE 100-step run.
    measure(ctx, o, xyz, pn, "aos4", "warp", 0, "force_kernel_warp_unroll2_double4", false);
    measure(ctx, o, xyz, pn, "aos3", "warp", 0, "force_kernel_warp_unroll2_double3", true);
  } else if (o.all) {
    for (const char* lay : {"aos3", "aos4", "soa"}) {
      for (int g : {1, 4, 8, 16, 32}) {
        const std::string name = std::string("force_gather_g") + std::to_string(g) + "_" + lay;
        measure(ctx, o, xyz, pn, lay, "subwarp", g, name.c_str(), false);
      }
      measure(ctx, o, xyz, pn, lay, "tile", 8, (std::string("force_tile_tma_g8_") + lay).c_str(), false);
      measure(ctx, o, xyz, pn, lay, "n3", 8, (std::string("force_newton3_g8_") + lay).c_str(), false);
    }
  } else {
    const std::string name = "force_" + o.variant + "_" + o.layout;
    measure(ctx, o, xyz, pn, o.layout, o.variant, o.group, name.c_str(), o.print);
  }
  lj_ctx_destroy(ctx);
  return 0;
}
