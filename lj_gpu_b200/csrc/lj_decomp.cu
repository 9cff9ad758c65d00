// lj_decomp.cu -- z-slab decomposition over the GPUs of one box, ONE process, behind the C ABI.
//
// No counterpart in the reference (single GPU: SURVEY 0.9).  The path shards with one exchange step
// (SURVEY 8e): a gather kernel only writes p[i] of owned particles, so per step a slab needs the
// current POSITIONS of the ghost particles within the search length of its faces and nothing flows
// back.  The caller hands over the particles in an order in which every slab is a contiguous index
// range and so are the rows its neighbours need (the reference's generator emits lattice layers
// z-outermost, cuda/force_cuda.cu:68-77; lj_decomp_plan_fcc computes the ranges): no packing.
//
//   slab g:  q_local = [ owned | ghosts from slab g-1 (its top rows) | ghosts from slab g+1 (its bottom rows) ]
//
// One lj_ctx per device, peer access between neighbours.  Per step and slab:
//   compute stream:  publish (lj_flag_set: "my q of step k is final")
//   comm stream:     lj_halo_pull_sync -- copies both ghost segments straight out of the neighbours'
//                    memory; each segment waits ON THE DEVICE for the owner's flag of step k and then
//                    tells the owner it is done (a counter in the owner's memory)
//   compute stream:  lj_force_step_part INTERIOR (tiles that read no ghost) while the halo flies,
//                    wait for the halo event, lj_force_step_part BOUNDARY
// and before q is overwritten (lj_decomp_md: drift) the owner waits for both neighbours' counters.
// The Python twin (lj_gpu_b200/decomp.py, one PROCESS per GPU for torchrun / bench.py) uses the same
// entry points with CUDA-IPC mappings instead of peer pointers.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "lj_common.cuh"

struct lj_decomp {
  int n = 0;
  std::vector<int> dev;
  std::vector<lj_ctx*> ctx;
  std::vector<int64_t> lo, hi, n_lo, n_hi;   // owned global range, ghosts received from below / above
  std::vector<double*> q, p;                  // [n_local][4] doubles (AOS_D4)
  std::vector<int32_t*> nop, list;
  std::vector<void*> ptr;
  std::vector<int64_t> cap, pairs;
  std::vector<int32_t*> flags;                // {READY, PULLED_BY_BELOW, PULLED_BY_ABOVE, pad}
  std::vector<cudaStream_t> comm;
  std::vector<cudaEvent_t> ev_q, ev_halo;
  std::vector<uint64_t> token;
  std::vector<int> ptr64;
  int64_t pn = 0;
  double search_len = 3.3, cl2 = 9.0, dt = 0.001;
  int precision = LJ_PREC_FP64, list_flags = LJ_LIST_TILES;
  int stepno = 0;
  std::string err;
};

namespace {

enum { READY = 0, PULLED_BY_BELOW = 1, PULLED_BY_ABOVE = 2 };

int fail(lj_decomp* d, int rc, const char* what, const char* detail) {
  if (d) d->err = std::string(what) + ": " + (detail ? detail : "");
  return rc;
}
#define D_CUDA(d, call)                                                               \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) return fail((d), LJ_ERR_CUDA, #call, cudaGetErrorString(e__)); \
  } while (0)
#define D_LJ(d, g, call)                                                              \
  do {                                                                                \
    int rc__ = (call);                                                                \
    if (rc__) return fail((d), rc__, #call, lj_last_error_string((d)->ctx[g]));        \
  } while (0)

int64_t n_own(const lj_decomp* d, int g) { return d->hi[g] - d->lo[g]; }
int64_t n_local(const lj_decomp* d, int g) { return n_own(d, g) + d->n_lo[g] + d->n_hi[g]; }

lj_list_args list_args(lj_decomp* d, int g) {
  lj_list_args a{};
  a.q = d->q[g]; a.pn = n_local(d, g); a.layout = LJ_AOS_D4; a.half = 0; a.search_len = d->search_len;
  a.number_of_partners = d->nop[g]; a.pointer = d->ptr[g]; a.sorted_list = d->list[g]; a.capacity = d->cap[g];
  a.pointer64 = d->ptr64[g]; a.flags = d->list_flags;
  a.row_begin = 0; a.row_end = n_own(d, g);
  return a;
}

lj_force_args force_args(lj_decomp* d, int g, double dt) {
  lj_force_args a{};
  a.q = d->q[g]; a.p = d->p[g]; a.pn = n_local(d, g); a.dt = dt; a.cl2 = d->cl2;
  a.list = d->list[g]; a.number_of_partners = d->nop[g]; a.pointer = d->ptr[g];
  a.layout = LJ_AOS_D4; a.list_layout = LJ_LIST_CSR; a.variant = LJ_VARIANT_AUTO; a.precision = d->precision;
  a.pointer64 = d->ptr64[g]; a.row_begin = 0; a.row_end = n_own(d, g); a.list_entries = d->cap[g];
  a.mirror_token = d->token[g];
  return a;
}

int rebuild_one(lj_decomp* d, int g, bool first) {
  lj_list_args a = list_args(d, g);
  if (first) {  // capacity protocol: a sizing call, then the real one
    a.sorted_list = nullptr; a.capacity = 0;
    int64_t np = 0;
    int rc = lj_build_list(d->ctx[g], &a, &np, nullptr);
    if (rc != LJ_OK && rc != LJ_ERR_CAPACITY) return fail(d, rc, "lj_build_list", lj_last_error_string(d->ctx[g]));
    d->cap[g] = np + np / 64 + 1024;
    D_LJ(d, g, lj_dev_alloc(d->ctx[g], sizeof(int32_t) * (size_t)d->cap[g], (void**)&d->list[g], nullptr));
    a = list_args(d, g);
    D_LJ(d, g, lj_build_list(d->ctx[g], &a, &np, nullptr));
    d->pairs[g] = np;
  } else {
    D_LJ(d, g, lj_build_list(d->ctx[g], &a, nullptr, nullptr));
  }
  d->token[g] = lj_list_mirror_token(d->ctx[g]);
  return LJ_OK;
}

// publish + halo of step d->stepno for every slab (all asynchronous)
int start_step(lj_decomp* d) {
  d->stepno++;
  for (int g = 0; g < d->n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    D_LJ(d, g, lj_flag_set(d->ctx[g], d->flags[g] + READY, d->stepno, nullptr));
    // the pulls of this step overwrite the ghost rows the previous step's boundary tiles read
    D_CUDA(d, cudaEventRecord(d->ev_q[g], nullptr));
  }
  for (int g = 0; g < d->n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    D_CUDA(d, cudaStreamWaitEvent(d->comm[g], d->ev_q[g], 0));
    lj_halo_seg segs[2];
    int ns = 0;
    const int64_t own = n_own(d, g);
    if (d->n_lo[g] > 0) {  // the top rows of the slab below; seen from there I am the neighbour ABOVE
      lj_halo_seg& s = segs[ns++];
      s.local_dst = d->q[g] + 4 * own;
      s.peer_src = d->q[g - 1] + 4 * (n_own(d, g - 1) - d->n_lo[g]);
      s.bytes = (size_t)d->n_lo[g] * 32;
      s.wait_flag = d->flags[g - 1] + READY; s.wait_value = d->stepno;
      s.done_flag = d->flags[g - 1] + PULLED_BY_ABOVE; s.done_value = d->stepno;
    }
    if (d->n_hi[g] > 0) {  // the bottom rows of the slab above
      lj_halo_seg& s = segs[ns++];
      s.local_dst = d->q[g] + 4 * (own + d->n_lo[g]);
      s.peer_src = d->q[g + 1];
      s.bytes = (size_t)d->n_hi[g] * 32;
      s.wait_flag = d->flags[g + 1] + READY; s.wait_value = d->stepno;
      s.done_flag = d->flags[g + 1] + PULLED_BY_BELOW; s.done_value = d->stepno;
    }
    if (ns) D_LJ(d, g, lj_halo_pull_sync(d->ctx[g], segs, ns, d->comm[g]));
    D_CUDA(d, cudaEventRecord(d->ev_halo[g], d->comm[g]));
  }
  return LJ_OK;
}

int force_all(lj_decomp* d, bool rebuild, bool overlap, double dt) {
  for (int g = 0; g < d->n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    const bool parts = overlap && !rebuild && d->token[g] != 0 && (d->n_lo[g] > 0 || d->n_hi[g] > 0);
    if (rebuild) {
      D_CUDA(d, cudaStreamWaitEvent(nullptr, d->ev_halo[g], 0));
      int rc = rebuild_one(d, g, false);
      if (rc) return rc;
    }
    lj_force_args a = force_args(d, g, dt);
    if (parts) {
      D_LJ(d, g, lj_force_step_part(d->ctx[g], &a, LJ_PART_INTERIOR, nullptr));
      D_CUDA(d, cudaStreamWaitEvent(nullptr, d->ev_halo[g], 0));
      D_LJ(d, g, lj_force_step_part(d->ctx[g], &a, LJ_PART_BOUNDARY, nullptr));
    } else {
      if (!rebuild) D_CUDA(d, cudaStreamWaitEvent(nullptr, d->ev_halo[g], 0));
      D_LJ(d, g, lj_force_step(d->ctx[g], &a, nullptr));
    }
  }
  return LJ_OK;
}

}  // namespace

extern "C" {

int lj_decomp_plan_fcc(double density, double L, int32_t ngpus, double search_len, int64_t* slab_begin,
                       int64_t* halo_rows) {
  if (!slab_begin || !halo_rows || ngpus < 1 || density <= 0.0 || L <= 0.0) return LJ_ERR_BAD_ARG;
  const double s = 1.0 / std::pow(density * 0.25, 1.0 / 3.0);
  const int64_t n = (int64_t)(L / s);
  if (ngpus > n) return LJ_ERR_BAD_ARG;
  const int64_t layer = 4 * n * n;
  // layer iz holds z in [iz*s, iz*s + s/2 + 0.1): layers that can hold a neighbour of a slab face
  const int64_t halo = (int64_t)std::ceil((search_len + 0.5 * s + 0.1) / s);
  for (int g = 0; g <= ngpus; g++) slab_begin[g] = (n * g) / ngpus * layer;
  for (int g = 0; g < ngpus; g++)
    if (ngpus > 1 && (slab_begin[g + 1] - slab_begin[g]) / layer < halo) return LJ_ERR_BAD_ARG;  // ghosts would live two slabs away
  *halo_rows = halo * layer;
  return LJ_OK;
}

int lj_decomp_create(lj_decomp** out, const lj_decomp_args* a) {
  if (!out || !a || a->ngpus < 1 || !a->q_xyz_host || !a->slab_begin || a->pn <= 0) return LJ_ERR_BAD_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return LJ_ERR_NO_DEVICE;
  lj_decomp* d = new lj_decomp;
  *out = d;  // returned even on failure so that lj_decomp_last_error can tell why
  const int n = d->n = a->ngpus;
  d->pn = a->pn;
  d->search_len = a->search_len > 0 ? a->search_len : 3.3;
  d->cl2 = (a->cutoff > 0 ? a->cutoff : 3.0) * (a->cutoff > 0 ? a->cutoff : 3.0);
  d->dt = a->dt > 0 ? a->dt : 0.001;
  d->precision = a->precision;
  d->list_flags = a->list_flags ? a->list_flags : (a->precision == LJ_PREC_MIXED ? (LJ_LIST_TILES | LJ_LIST_TILES_WIDE) : LJ_LIST_TILES);
  d->dev.resize(n); d->ctx.assign(n, nullptr); d->lo.resize(n); d->hi.resize(n); d->n_lo.assign(n, 0); d->n_hi.assign(n, 0);
  d->q.assign(n, nullptr); d->p.assign(n, nullptr); d->nop.assign(n, nullptr); d->list.assign(n, nullptr);
  d->ptr.assign(n, nullptr); d->cap.assign(n, 0); d->pairs.assign(n, 0); d->flags.assign(n, nullptr);
  d->comm.assign(n, nullptr); d->ev_q.assign(n, nullptr); d->ev_halo.assign(n, nullptr); d->token.assign(n, 0);
  d->ptr64.assign(n, 0);
  for (int g = 0; g < n; g++) {
    d->dev[g] = a->devices ? a->devices[g] : g;
    if (d->dev[g] < 0 || d->dev[g] >= count) return fail(d, LJ_ERR_BAD_ARG, "lj_decomp_create", "device ordinal out of range");
    d->lo[g] = a->slab_begin[g]; d->hi[g] = a->slab_begin[g + 1];
    if (d->hi[g] <= d->lo[g] || d->hi[g] > a->pn) return fail(d, LJ_ERR_BAD_ARG, "lj_decomp_create", "slab_begin must ascend within [0, pn]");
  }
  if (d->lo[0] != 0 || d->hi[n - 1] != a->pn) return fail(d, LJ_ERR_BAD_ARG, "lj_decomp_create", "the slabs must cover [0, pn)");
  for (int g = 0; g < n; g++) {
    if (g > 0) d->n_lo[g] = a->halo_rows < d->hi[g - 1] - d->lo[g - 1] ? a->halo_rows : -1;
    if (g < n - 1) d->n_hi[g] = a->halo_rows < d->hi[g + 1] - d->lo[g + 1] ? a->halo_rows : -1;
    if (d->n_lo[g] < 0 || d->n_hi[g] < 0)
      return fail(d, LJ_ERR_BAD_ARG, "lj_decomp_create", "a slab is thinner than the halo: its ghosts would live two slabs away");
  }
  std::vector<double> stage;
  for (int g = 0; g < n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    int rc = lj_ctx_create(&d->ctx[g], d->dev[g]);
    if (rc) return fail(d, rc, "lj_ctx_create", "");
    const int64_t nl = n_local(d, g), own = n_own(d, g);
    D_CUDA(d, cudaMalloc((void**)&d->q[g], (size_t)nl * 32));   // plain cudaMalloc: peers address it directly
    D_CUDA(d, cudaMalloc((void**)&d->p[g], (size_t)nl * 32));
    D_CUDA(d, cudaMalloc((void**)&d->flags[g], 4 * sizeof(int32_t)));
    D_CUDA(d, cudaMemset(d->flags[g], 0, 4 * sizeof(int32_t)));
    D_CUDA(d, cudaMemset(d->p[g], 0, (size_t)nl * 32));
    D_CUDA(d, cudaMalloc((void**)&d->nop[g], (size_t)nl * sizeof(int32_t)));
    d->ptr64[g] = own * 160 > (1LL << 31);   // int32 pointer[] holds 2^32 - 1 list entries
    D_CUDA(d, cudaMalloc(&d->ptr[g], (size_t)nl * (d->ptr64[g] ? 8 : 4)));
    D_CUDA(d, cudaStreamCreateWithFlags(&d->comm[g], cudaStreamNonBlocking));
    D_CUDA(d, cudaEventCreateWithFlags(&d->ev_q[g], cudaEventDisableTiming));
    D_CUDA(d, cudaEventCreateWithFlags(&d->ev_halo[g], cudaEventDisableTiming));
    // local positions: [owned | top rows of the slab below | bottom rows of the slab above], xyz -> double4
    stage.assign((size_t)nl * 4, 0.0);
    auto put = [&](int64_t dst, int64_t src, int64_t cnt) {
      for (int64_t i = 0; i < cnt; i++)
        for (int c = 0; c < 3; c++) stage[(size_t)(dst + i) * 4 + c] = a->q_xyz_host[(size_t)(src + i) * 3 + c];
    };
    put(0, d->lo[g], own);
    if (d->n_lo[g]) put(own, d->lo[g] - d->n_lo[g], d->n_lo[g]);
    if (d->n_hi[g]) put(own + d->n_lo[g], d->hi[g], d->n_hi[g]);
    D_CUDA(d, cudaMemcpy(d->q[g], stage.data(), (size_t)nl * 32, cudaMemcpyHostToDevice));
  }
  for (int g = 0; g < n; g++) {  // neighbours read each other's q and flags
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    for (int o : {g - 1, g + 1}) {
      if (o < 0 || o >= n || d->dev[o] == d->dev[g]) continue;
      int can = 0;
      D_CUDA(d, cudaDeviceCanAccessPeer(&can, d->dev[g], d->dev[o]));
      if (!can) return fail(d, LJ_ERR_CUDA, "lj_decomp_create", "no peer access between neighbouring devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(d->dev[o], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(d, LJ_ERR_CUDA, "cudaDeviceEnablePeerAccess", cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  for (int g = 0; g < n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    int rc = rebuild_one(d, g, true);
    if (rc) return rc;
  }
  return lj_decomp_sync(d);
}

int lj_decomp_step(lj_decomp* d, int32_t nsteps, int32_t rebuild_every, int32_t overlap) {
  if (!d || nsteps < 0) return LJ_ERR_BAD_ARG;
  for (int k = 0; k < nsteps; k++) {
    int rc = start_step(d);
    if (rc) return rc;
    // static positions: a rebuild gives the same list, the cadence measures its cost (as in bench.py)
    rc = force_all(d, rebuild_every > 0 && (d->stepno - 1) % rebuild_every == 0 && d->stepno > 1, overlap != 0, d->dt);
    if (rc) return rc;
  }
  return LJ_OK;
}

int lj_decomp_md(lj_decomp* d, int32_t nsteps, int32_t rebuild_every, int32_t overlap) {
  if (!d || nsteps < 0) return LJ_ERR_BAD_ARG;
  for (int k = 0; k < nsteps; k++) {
    int rc = start_step(d);
    if (rc) return rc;
    rc = force_all(d, rebuild_every > 0 && k > 0 && k % rebuild_every == 0, overlap != 0, d->dt);
    if (rc) return rc;
    for (int g = 0; g < d->n; g++) {  // drift of the owned particles, once both neighbours have read this step's q
      D_CUDA(d, cudaSetDevice(d->dev[g]));
      if (g > 0) D_LJ(d, g, lj_flag_wait(d->ctx[g], d->flags[g] + PULLED_BY_BELOW, d->stepno, nullptr));
      if (g < d->n - 1) D_LJ(d, g, lj_flag_wait(d->ctx[g], d->flags[g] + PULLED_BY_ABOVE, d->stepno, nullptr));
      D_LJ(d, g, lj_drift(d->ctx[g], d->q[g], d->p[g], n_own(d, g), LJ_AOS_D4, 0, d->dt, nullptr));
    }
  }
  return LJ_OK;
}

int lj_decomp_rebuild(lj_decomp* d) {
  if (!d) return LJ_ERR_BAD_ARG;
  for (int g = 0; g < d->n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    int rc = rebuild_one(d, g, false);
    if (rc) return rc;
  }
  return LJ_OK;
}

int lj_decomp_sync(lj_decomp* d) {
  if (!d) return LJ_ERR_BAD_ARG;
  for (int g = 0; g < d->n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    D_CUDA(d, cudaStreamSynchronize(d->comm[g]));
    D_CUDA(d, cudaDeviceSynchronize());
  }
  return LJ_OK;
}

int lj_decomp_gather(lj_decomp* d, double* p_xyz_host, double* q_xyz_host) {
  if (!d) return LJ_ERR_BAD_ARG;
  int rc = lj_decomp_sync(d);
  if (rc) return rc;
  std::vector<double> stage;
  for (int g = 0; g < d->n; g++) {
    D_CUDA(d, cudaSetDevice(d->dev[g]));
    const int64_t own = n_own(d, g);
    stage.resize((size_t)own * 4);
    for (int which = 0; which < 2; which++) {
      double* out = which ? q_xyz_host : p_xyz_host;
      if (!out) continue;
      D_CUDA(d, cudaMemcpy(stage.data(), which ? d->q[g] : d->p[g], (size_t)own * 32, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < own; i++)
        for (int c = 0; c < 3; c++) out[(size_t)(d->lo[g] + i) * 3 + c] = stage[(size_t)i * 4 + c];
    }
  }
  return LJ_OK;
}

int64_t lj_decomp_pairs(lj_decomp* d) {
  int64_t t = 0;
  if (d) for (int64_t v : d->pairs) t += v;
  return t;
}

int64_t lj_decomp_launch_count(lj_decomp* d) {
  int64_t t = 0;
  if (d) for (lj_ctx* c : d->ctx) if (c) t += lj_launch_count(c);
  return t;
}

const char* lj_decomp_last_error(lj_decomp* d) { return d ? d->err.c_str() : "null decomposition"; }

int lj_decomp_destroy(lj_decomp* d) {
  if (!d) return LJ_ERR_BAD_ARG;
  for (int g = 0; g < d->n; g++) {
    cudaSetDevice(d->dev[g]);
    cudaDeviceSynchronize();
    if (d->list[g] && d->ctx[g]) lj_dev_free(d->ctx[g], d->list[g], nullptr);
    for (void* v : {(void*)d->q[g], (void*)d->p[g], (void*)d->flags[g], (void*)d->nop[g], d->ptr[g]})
      if (v) cudaFree(v);
    if (d->comm[g]) cudaStreamDestroy(d->comm[g]);
    if (d->ev_q[g]) cudaEventDestroy(d->ev_q[g]);
    if (d->ev_halo[g]) cudaEventDestroy(d->ev_halo[g]);
    if (d->ctx[g]) lj_ctx_destroy(d->ctx[g]);
  }
  delete d;
  return LJ_OK;
}

}  // extern "C"
