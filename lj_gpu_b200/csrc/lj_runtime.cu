// lj_runtime.cu -- context, Blackwell memory layer and the measure() call of the C ABI.
//
// Replaces cuda/cuda_ptr.cuh (paired cudaMalloc + cudaMallocHost sized by a static maximum,
// blocking cudaMemcpy, thrust::fill) with: a per-context stream-ordered pool
// (cudaMallocAsync, release threshold = keep everything), pinned host mirrors, async copies
// on the caller's stream, and a double-buffered pinned staging ring for pageable memory.
#include <chrono>
#include <cstring>

#include "lj_common.cuh"

int lj_set_error(lj_ctx* ctx, int status, const char* what, const char* detail) {
  if (ctx) {
    ctx->err = std::string(what ? what : "") + ": " + (detail ? detail : "");
  }
  return status;
}

namespace {
constexpr size_t kRingBytes = 32u << 20;  // 2 x 32 MiB pinned staging

__global__ void k_fill32(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// 16-byte grid-stride copy; used on peer-mapped sources (P2P loads over NVLink)
__global__ void __launch_bounds__(256) k_copy16(int4* __restrict__ dst, const int4* __restrict__ src, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

// See LJ_FUNC_SMEM: opt `kern` in to the most dynamic shared memory the device allows next to the kernel's
// static shared memory (never to less than another context on the same device asked for).
int lj_func_smem(lj_ctx* ctx, const void* kern, size_t bytes) {
  size_t& have = ctx->func_smem[kern];
  if (bytes <= have) return LJ_OK;
  cudaFuncAttributes fa{};
  LJ_CUDA(ctx, cudaFuncGetAttributes(&fa, kern));
  const size_t most = ctx->smem_optin > fa.sharedSizeBytes ? ctx->smem_optin - fa.sharedSizeBytes : 0;
  LJ_REQUIRE(ctx, bytes <= most, "kernel needs more shared memory than the device offers");
  LJ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
  have = most;
  return LJ_OK;
}

// ---- lj_kernel_timing: event pairs around the dominant force kernel (see lj_b200.h) ----
static void kt_fold(lj_ctx* ctx, int pair) {  // waits for the pair's second event
  float ms = 0.f;
  if (cudaEventSynchronize(ctx->kt_ev[2 * pair + 1]) == cudaSuccess &&
      cudaEventElapsedTime(&ms, ctx->kt_ev[2 * pair], ctx->kt_ev[2 * pair + 1]) == cudaSuccess) {
    ctx->kt_ms += ms;
    ctx->kt_launches++;
  } else {
    cudaGetLastError();
  }
}
static void kt_fold_all(lj_ctx* ctx) {
  const int n = 64;
  for (int k = ctx->kt_pending; k > 0; k--) kt_fold(ctx, ((ctx->kt_head - k) % n + n) % n);
  ctx->kt_pending = 0;
}
// called by the launcher of the dominant kernel: the event to record BEFORE (first = true) / AFTER the launch,
// or nullptr when timing is off or the stream is being captured
cudaEvent_t lj_kernel_timing_event(lj_ctx* ctx, cudaStream_t st, bool first) {
  if (!ctx->kt_on) return nullptr;
  const int n = 64;
  if (first) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    if (ctx->kt_pending == n) { kt_fold(ctx, ctx->kt_head); ctx->kt_pending--; }  // the oldest pair is reused
    for (int e = 0; e < 2; e++)
      if (!ctx->kt_ev[2 * ctx->kt_head + e] && cudaEventCreate(&ctx->kt_ev[2 * ctx->kt_head + e]) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
    return ctx->kt_ev[2 * ctx->kt_head];
  }
  cudaEvent_t e = ctx->kt_ev[2 * ctx->kt_head + 1];
  ctx->kt_head = (ctx->kt_head + 1) % n;
  ctx->kt_pending++;
  return e;
}

extern "C" {

const char* lj_status_string(int s) {
  switch (s) {
    case LJ_OK: return "ok";
    case LJ_ERR_CUDA: return "CUDA error";
    case LJ_ERR_BAD_ARG: return "bad argument";
    case LJ_ERR_CAPACITY: return "capacity overflow";
    case LJ_ERR_OVERFLOW32: return "32-bit offset overflow";
    case LJ_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case LJ_ERR_INVALID_LIST: return "invalid neighbour list";
  }
  return "unknown status";
}

int lj_ctx_create(lj_ctx** out, int device) {
  if (!out) return LJ_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return LJ_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return LJ_ERR_BAD_ARG;
  lj_ctx* ctx = new lj_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return LJ_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
  }
  cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  // private pool: never hand memory back to the OS between steps (180 GB HBM, one tenant)
  cudaMemPoolProps pp{};
  pp.allocType = cudaMemAllocationTypePinned;
  pp.handleTypes = cudaMemHandleTypeNone;
  pp.location.type = cudaMemLocationTypeDevice;
  pp.location.id = device;
  if (cudaMemPoolCreate(&ctx->pool, &pp) != cudaSuccess) {
    cudaGetLastError();
    cudaDeviceGetDefaultMemPool(&ctx->pool, device);
  }
  uint64_t keep = ~0ull;
  cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
  for (int k = 0; k < 2; k++) cudaEventCreateWithFlags(&ctx->ring_ev[k], cudaEventDisableTiming);
  if (cudaGetLastError() != cudaSuccess) { /* non-fatal: reported on first use */ }
  *out = ctx;
  return LJ_OK;
}

int lj_ctx_destroy(lj_ctx* ctx) {
  if (!ctx) return LJ_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  void* frees[] = {ctx->bbox, ctx->grid, ctx->totals, ctx->cell_of, ctx->cell_slot, ctx->cell_count,
                   ctx->cell_start, ctx->sorted_pos, ctx->sorted_tmp, ctx->scan_tmp, ctx->q32,
                   ctx->cl_list, ctx->cl_ptr, ctx->cl_cnt, ctx->tl_geom, ctx->tl_order, ctx->tl_cnt,
                   ctx->tl_units, ctx->tl_off, ctx->tl_qs, ctx->tl_qfx, ctx->soa6_q, ctx->soa6_p, ctx->tl_cell_start, ctx->tl_list, ctx->tl_tab, ctx->tl_ttab, ctx->tl_cols, ctx->tl_meta, ctx->tl_zflag, ctx->tl_cols_sel, ctx->tl_rowflag, ctx->tl_slot_of};
  for (void* f : frees)
    if (f) cudaFreeAsync(f, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->totals_host) cudaFreeHost(ctx->totals_host);
  if (ctx->tl_geom_host) cudaFreeHost(ctx->tl_geom_host);
  for (cudaEvent_t e : ctx->kt_ev)
    if (e) cudaEventDestroy(e);
  for (int k = 0; k < 2; k++) {
    if (ctx->ring[k]) cudaFreeHost(ctx->ring[k]);
    if (ctx->ring_ev[k]) cudaEventDestroy(ctx->ring_ev[k]);
  }
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  cudaMemPool_t def = nullptr;
  cudaDeviceGetDefaultMemPool(&def, ctx->device);
  if (ctx->pool && ctx->pool != def) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
  return LJ_OK;
}

int lj_sync(lj_ctx* ctx, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (stream) {
    LJ_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
  } else {
    LJ_CUDA(ctx, cudaStreamSynchronize(nullptr));  // legacy default stream
    LJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LJ_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  }
  return LJ_OK;
}

int lj_list_invalidate(lj_ctx* ctx) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  ctx->cl_valid = false;
  ctx->tl_valid = false;
  ctx->graph_loop = -1;  // a cached CUDA graph may have captured the kernel that ran on the mirror
  return LJ_OK;
}

const char* lj_last_error_string(lj_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t lj_launch_count(lj_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lj_kernel_timing(lj_ctx* ctx, int enable) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (enable) { kt_fold_all(ctx); ctx->kt_ms = 0.0; ctx->kt_launches = 0; }
  ctx->kt_on = enable != 0;
  return LJ_OK;
}
int lj_kernel_timing_read(lj_ctx* ctx, double* total_ms_out, int64_t* launches_out) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  kt_fold_all(ctx);
  if (total_ms_out) *total_ms_out = ctx->kt_ms;
  if (launches_out) *launches_out = ctx->kt_launches;
  return LJ_OK;
}
void* lj_ctx_stream(lj_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int lj_device_sm_count(lj_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// ------------------------------------------------------------------ memory layer ------
int lj_dev_alloc(lj_ctx* ctx, size_t bytes, void** out, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx || !out) return LJ_ERR_BAD_ARG;
  *out = nullptr;
  if (bytes == 0) return LJ_OK;
  LJ_CUDA(ctx, cudaMallocAsync(out, bytes, ctx->pool, lj_stream(ctx, stream)));
  return LJ_OK;
}

int lj_dev_free(lj_ctx* ctx, void* ptr, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (ptr) LJ_CUDA(ctx, cudaFreeAsync(ptr, lj_stream(ctx, stream)));
  return LJ_OK;
}

int lj_buf_allocate(lj_ctx* ctx, size_t bytes, lj_buf* out, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx || !out) return LJ_ERR_BAD_ARG;
  out->host = out->dev = nullptr;
  out->bytes = bytes;
  if (bytes == 0) return LJ_OK;
  LJ_CUDA(ctx, cudaMallocAsync(&out->dev, bytes, ctx->pool, lj_stream(ctx, stream)));
  cudaError_t e = cudaHostAlloc(&out->host, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaFreeAsync(out->dev, lj_stream(ctx, stream));
    out->dev = nullptr;
    return lj_set_error(ctx, LJ_ERR_CUDA, "cudaHostAlloc", cudaGetErrorString(e));
  }
  return LJ_OK;
}

int lj_buf_deallocate(lj_ctx* ctx, lj_buf* buf, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx || !buf) return LJ_ERR_BAD_ARG;
  if (buf->dev) LJ_CUDA(ctx, cudaFreeAsync(buf->dev, lj_stream(ctx, stream)));
  if (buf->host) {
    // the host mirror may still be the target of an in-flight copy on this stream
    LJ_CUDA(ctx, cudaStreamSynchronize(lj_stream(ctx, stream)));
    LJ_CUDA(ctx, cudaFreeHost(buf->host));
  }
  buf->host = buf->dev = nullptr;
  buf->bytes = 0;
  return LJ_OK;
}

int lj_buf_host2dev(lj_ctx* ctx, const lj_buf* buf, size_t beg, size_t count, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx || !buf) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, beg + count <= buf->bytes, "lj_buf_host2dev: range exceeds the buffer");
  if (count == 0) return LJ_OK;
  LJ_CUDA(ctx, cudaMemcpyAsync((char*)buf->dev + beg, (const char*)buf->host + beg, count,
                               cudaMemcpyHostToDevice, lj_stream(ctx, stream)));
  return LJ_OK;
}

int lj_buf_dev2host(lj_ctx* ctx, const lj_buf* buf, size_t beg, size_t count, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx || !buf) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, beg + count <= buf->bytes, "lj_buf_dev2host: range exceeds the buffer");
  if (count == 0) return LJ_OK;
  LJ_CUDA(ctx, cudaMemcpyAsync((char*)buf->host + beg, (const char*)buf->dev + beg, count,
                               cudaMemcpyDeviceToHost, lj_stream(ctx, stream)));
  return LJ_OK;
}

int lj_buf_set_val32(lj_ctx* ctx, const lj_buf* buf, size_t beg, size_t count, uint32_t value,
                     void* stream) {
  LJ_ENTER(ctx);
  if (!ctx || !buf) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, (beg + count) * 4 <= buf->bytes, "lj_buf_set_val32: range exceeds the buffer");
  if (count == 0) return LJ_OK;
  uint32_t* h = (uint32_t*)buf->host + beg;
  for (size_t i = 0; i < count; i++) h[i] = value;
  size_t blocks = (count + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 8) blocks = (size_t)ctx->sm_count * 8;
  k_fill32<<<(unsigned)blocks, 256, 0, lj_stream(ctx, stream)>>>((uint32_t*)buf->dev + beg, count, value);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

static int ring_init(lj_ctx* ctx) {
  if (ctx->ring[0]) return LJ_OK;
  for (int k = 0; k < 2; k++) LJ_CUDA(ctx, cudaHostAlloc(&ctx->ring[k], kRingBytes, cudaHostAllocDefault));
  ctx->ring_bytes = kRingBytes;
  return LJ_OK;
}

int lj_upload(lj_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (bytes == 0) return LJ_OK;
  LJ_REQUIRE(ctx, dev_dst && host_src, "lj_upload: null pointer");
  cudaStream_t st = lj_stream(ctx, stream);
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, host_src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
    LJ_CUDA(ctx, cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, st));  // already pinned
    return LJ_OK;
  }
  cudaGetLastError();
  int rc = ring_init(ctx);
  if (rc) return rc;
  size_t off = 0;
  int k = 0;
  while (off < bytes) {
    const size_t n = bytes - off < ctx->ring_bytes ? bytes - off : ctx->ring_bytes;
    LJ_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[k]));  // previous DMA out of this half done
    memcpy(ctx->ring[k], (const char*)host_src + off, n);  // overlaps the other half's DMA
    LJ_CUDA(ctx, cudaMemcpyAsync((char*)dev_dst + off, ctx->ring[k], n, cudaMemcpyHostToDevice, st));
    LJ_CUDA(ctx, cudaEventRecord(ctx->ring_ev[k], st));
    off += n;
    k ^= 1;
  }
  return LJ_OK;
}

int lj_download(lj_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (bytes == 0) return LJ_OK;
  LJ_REQUIRE(ctx, host_dst && dev_src, "lj_download: null pointer");
  cudaStream_t st = lj_stream(ctx, stream);
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, host_dst) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
    LJ_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, st));
    return LJ_OK;
  }
  cudaGetLastError();
  int rc = ring_init(ctx);
  if (rc) return rc;
  // chunk k is copied out of the ring while chunk k+1 is in flight
  size_t off = 0, prev_off = 0, prev_n = 0;
  int k = 0;
  while (off < bytes) {
    const size_t n = bytes - off < ctx->ring_bytes ? bytes - off : ctx->ring_bytes;
    LJ_CUDA(ctx, cudaMemcpyAsync(ctx->ring[k], (const char*)dev_src + off, n, cudaMemcpyDeviceToHost, st));
    LJ_CUDA(ctx, cudaEventRecord(ctx->ring_ev[k], st));
    if (prev_n) {
      LJ_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[k ^ 1]));
      memcpy((char*)host_dst + prev_off, ctx->ring[k ^ 1], prev_n);
    }
    prev_off = off; prev_n = n;
    off += n;
    k ^= 1;
  }
  LJ_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[k ^ 1]));
  memcpy((char*)host_dst + prev_off, ctx->ring[k ^ 1], prev_n);
  return LJ_OK;
}

// ------------------------------------------------------------------ force entry points -
int lj_force_step(lj_ctx* ctx, const lj_force_args* args, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  return lj_force_launch(ctx, args, lj_stream(ctx, stream));
}

int lj_force_step_part(lj_ctx* ctx, const lj_force_args* args, int part, void* stream) {
  LJ_ENTER(ctx);
  return lj_force_launch(ctx, args, lj_stream(ctx, stream), part);
}

int lj_force_loop(lj_ctx* ctx, const lj_force_args* args, int loop, int use_graph, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, args != nullptr && loop >= 0, "lj_force_loop: bad arguments");
  cudaStream_t st = lj_stream(ctx, stream);
  if (!use_graph || loop < 2) {
    for (int s = 0; s < loop; s++) {
      int rc = lj_force_launch(ctx, args, st);
      if (rc) return rc;
    }
    return LJ_OK;
  }
  // capture once, replay: the 100-launch loop of measure() becomes one graph launch
  const bool same = ctx->graph_exec && ctx->graph_loop == loop &&
                    memcmp(&ctx->graph_args, args, sizeof(*args)) == 0;
  if (!same) {
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    // Validate the arguments and force the kernel's module to load OUTSIDE the capture with
    // a numerically neutral launch: a few rows, dt = 0 (p += 0).
    lj_force_args warm = *args;
    warm.dt = 0.0;
    warm.row_begin = 0;
    warm.row_end = args->pn < 32 ? args->pn : 32;
    if (args->row_begin || args->row_end) {
      warm.row_begin = args->row_begin;
      warm.row_end = args->row_begin + 32 < args->row_end ? args->row_begin + 32 : args->row_end;
    }
    {  // the cell-tile kernel only runs on the row range of its mirror: warm it with a full, neutral step
      int64_t r0 = args->row_begin, r1 = args->row_end;
      if (r0 == 0 && r1 == 0) r1 = args->pn;
      if ((args->variant == LJ_VARIANT_CELLTILE || args->variant == LJ_VARIANT_AUTO) &&
          lj_celltile_usable(ctx, args, r0, r1)) {
        warm.row_begin = args->row_begin;
        warm.row_end = args->row_end;
      }
    }
    int rc = lj_force_launch(ctx, &warm, st);
    if (rc) return rc;
    LJ_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t before = ctx->launches;
    cudaStream_t cap = ctx->copy_stream;  // never the stream the caller is working on
    LJ_CUDA(ctx, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    for (int s = 0; s < loop && rc == LJ_OK; s++) rc = lj_force_launch(ctx, args, cap);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cap, &graph);
    ctx->graph_step_launches = loop > 0 ? (ctx->launches - before) / loop : 0;
    ctx->launches = before;  // captured, not executed yet
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return lj_set_error(ctx, LJ_ERR_CUDA, "cudaStreamEndCapture", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      ctx->graph_exec = nullptr;
      return lj_set_error(ctx, LJ_ERR_CUDA, "cudaGraphInstantiate", cudaGetErrorString(e));
    }
    ctx->graph_args = *args;
    ctx->graph_loop = loop;
  }
  LJ_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, st));
  ctx->launches += ctx->graph_step_launches * loop;  // kernels the replay executes
  return LJ_OK;
}

// ------------------------------------------------------------------ six-array SoA ------
// openacc/force_oacc_soa.cpp:17-22 keeps qx,qy,qz,px,py,pz as six separate allocations.
static bool soa6_in_place(const double* x, const double* y, const double* z, int64_t pn, int64_t* stride) {
  const ptrdiff_t d1 = y - x, d2 = z - y;
  if (d1 != d2 || d1 < pn) return false;
  *stride = (int64_t)d1;
  return true;
}

static int soa6_reserve(lj_ctx* ctx, int64_t pn, cudaStream_t st) {
  if (pn <= ctx->soa6_cap) return LJ_OK;
  if (ctx->soa6_q) LJ_CUDA(ctx, cudaFreeAsync(ctx->soa6_q, st));
  if (ctx->soa6_p) LJ_CUDA(ctx, cudaFreeAsync(ctx->soa6_p, st));
  ctx->soa6_q = ctx->soa6_p = nullptr;
  const int64_t cap = (pn + 31) & ~(int64_t)31;
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->soa6_q, sizeof(double) * 3 * (size_t)cap, ctx->pool, st));
  LJ_CUDA(ctx, cudaMallocAsync((void**)&ctx->soa6_p, sizeof(double) * 3 * (size_t)cap, ctx->pool, st));
  ctx->soa6_cap = cap;
  ctx->graph_loop = -1;
  return LJ_OK;
}

static int soa6_gather(lj_ctx* ctx, double* dst, const double* x, const double* y, const double* z,
                       int64_t pn, cudaStream_t st) {
  const size_t b = sizeof(double) * (size_t)pn;
  const int64_t cap = ctx->soa6_cap;
  LJ_CUDA(ctx, cudaMemcpyAsync(dst, x, b, cudaMemcpyDeviceToDevice, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(dst + cap, y, b, cudaMemcpyDeviceToDevice, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(dst + 2 * cap, z, b, cudaMemcpyDeviceToDevice, st));
  return LJ_OK;
}

int lj_force_loop_soa6(lj_ctx* ctx, const double* qx, const double* qy, const double* qz, double* px,
                       double* py, double* pz, const lj_force_args* fa, int loop, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, fa && qx && qy && qz && px && py && pz && loop >= 0, "lj_force_loop_soa6: bad arguments");
  LJ_REQUIRE(ctx, fa->pn >= 0, "lj_force_loop_soa6: negative particle_number");
  cudaStream_t st = lj_stream(ctx, stream);
  const int64_t pn = fa->pn;
  if (pn == 0 || loop == 0) return LJ_OK;
  lj_force_args a = *fa;
  a.layout = LJ_SOA_D;
  int64_t sq = 0, sp = 0;
  const bool q_in_place = soa6_in_place(qx, qy, qz, pn, &sq);
  const bool p_in_place = soa6_in_place(px, py, pz, pn, &sp);
  if (q_in_place && p_in_place && sq == sp) {  // one block each, same spacing: nothing to copy
    a.q = qx; a.p = px; a.plane_stride = sq;
    return lj_force_loop(ctx, &a, loop, 0, stream);
  }
  int rc = soa6_reserve(ctx, pn, st);
  if (rc) return rc;
  if ((rc = soa6_gather(ctx, ctx->soa6_q, qx, qy, qz, pn, st))) return rc;
  if ((rc = soa6_gather(ctx, ctx->soa6_p, px, py, pz, pn, st))) return rc;
  a.q = ctx->soa6_q; a.p = ctx->soa6_p; a.plane_stride = ctx->soa6_cap;
  if ((rc = lj_force_loop(ctx, &a, loop, 0, stream))) return rc;
  const size_t b = sizeof(double) * (size_t)pn;
  LJ_CUDA(ctx, cudaMemcpyAsync(px, ctx->soa6_p, b, cudaMemcpyDeviceToDevice, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(py, ctx->soa6_p + ctx->soa6_cap, b, cudaMemcpyDeviceToDevice, st));
  LJ_CUDA(ctx, cudaMemcpyAsync(pz, ctx->soa6_p + 2 * ctx->soa6_cap, b, cudaMemcpyDeviceToDevice, st));
  return LJ_OK;
}

int lj_build_list_soa6(lj_ctx* ctx, const double* qx, const double* qy, const double* qz,
                       const lj_list_args* la, int64_t* number_of_pairs_out, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, la && qx && qy && qz, "lj_build_list_soa6: bad arguments");
  LJ_REQUIRE(ctx, la->pn >= 0 && la->pn < 2147483647LL, "lj_build_list_soa6: particle_number out of range");
  cudaStream_t st = lj_stream(ctx, stream);
  lj_list_args a = *la;
  a.layout = LJ_SOA_D;
  int64_t sq = 0;
  if (soa6_in_place(qx, qy, qz, a.pn, &sq)) {
    a.q = qx; a.plane_stride = sq;
  } else {
    int rc = soa6_reserve(ctx, a.pn, st);
    if (rc) return rc;
    if ((rc = soa6_gather(ctx, ctx->soa6_q, qx, qy, qz, a.pn, st))) return rc;
    a.q = ctx->soa6_q; a.plane_stride = ctx->soa6_cap;
  }
  // the cell-tile mirror is tied to the arrays of one call sequence; a gathered q block is
  // refreshed by every soa6 call, so the mirror stays valid for lj_force_loop_soa6
  return lj_build_list(ctx, &a, number_of_pairs_out, stream);
}

// ------------------------------------------------------------------ measure() ---------
static size_t vec_bytes(int layout) {
  return layout == LJ_AOS_D4 ? 32 : layout == LJ_AOS_D3 ? 24 : layout == LJ_AOS_F4 ? 16 : layout == LJ_AOS_F3 ? 12 : 8;
}

int lj_measure(lj_ctx* ctx, lj_measure_args* m) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, m && m->q_host && m->p_host && m->pn > 0 && m->loop >= 0, "lj_measure: bad arguments");
  LJ_REQUIRE(ctx, m->layout == LJ_AOS_D3 || m->layout == LJ_AOS_D4 || m->layout == LJ_SOA_D,
             "lj_measure: layout must be AOS_D3, AOS_D4 or SOA_D");
  cudaStream_t st = ctx->stream;
  const int64_t pn = m->pn;
  const size_t qbytes = m->layout == LJ_SOA_D ? (size_t)m->plane_stride * 3 * 8 : (size_t)pn * vec_bytes(m->layout);
  if (m->layout == LJ_SOA_D) LJ_REQUIRE(ctx, m->plane_stride >= pn, "lj_measure: plane_stride < pn");
  const bool own_list = m->list_host == nullptr;
  // validated before the first allocation: LJ_REQUIRE returns, it does not release anything
  LJ_REQUIRE(ctx, own_list || (m->number_of_partners_host && m->pointer_host && m->number_of_pairs_in >= 0),
             "lj_measure: incomplete host list");
  const double t_all0 = now_s();
  m->h2d_bytes = m->d2h_bytes = 0;
  m->list_builds = 0;

  void *q = nullptr, *p = nullptr;
  int32_t *list = nullptr, *nop = nullptr;
  void* ptr = nullptr;
  int rc = LJ_OK;
  int64_t capacity = 0, npairs = 0;
  int32_t max_np = 0;
  int ptr64 = own_list ? 1 : 0;  // int64 offsets whenever the library builds the list itself
  bool p_on_copy_stream = false;
#define M_TRY(x) do { rc = (x); if (rc) goto done; } while (0)
  M_TRY(lj_dev_alloc(ctx, qbytes, &q, st));
  M_TRY(lj_dev_alloc(ctx, qbytes, &p, st));
  M_TRY(lj_dev_alloc(ctx, sizeof(int32_t) * pn, (void**)&nop, st));
  M_TRY(lj_dev_alloc(ctx, (ptr64 ? 8 : 4) * (size_t)pn, &ptr, st));
  M_TRY(lj_upload(ctx, q, m->q_host, qbytes, st));
  // p is not needed before the first force step: its upload runs on the copy stream behind the list build
  {
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, m->p_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned && ctx->copy_stream && ctx->copy_stream != st) {
      if (!ctx->ev_copy) LJ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
      if ((rc = cudaEventRecord(ctx->ev_copy, st)) != cudaSuccess) goto cuda_fail;   // p was allocated on st
      if ((rc = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0)) != cudaSuccess) goto cuda_fail;
      if ((rc = cudaMemcpyAsync(p, m->p_host, qbytes, cudaMemcpyHostToDevice, ctx->copy_stream)) != cudaSuccess) goto cuda_fail;
      if ((rc = cudaEventRecord(ctx->ev_copy, ctx->copy_stream)) != cudaSuccess) goto cuda_fail;
      p_on_copy_stream = true;
    } else {
      M_TRY(lj_upload(ctx, p, m->p_host, qbytes, st));
    }
  }
  m->h2d_bytes += 2 * (int64_t)qbytes;

  {
    lj_list_args la{};
    la.q = q; la.pn = pn; la.layout = m->layout; la.half = m->half; la.plane_stride = m->plane_stride;
    la.search_len = m->search_len; la.number_of_partners = nop; la.pointer = ptr; la.pointer64 = ptr64;
    la.flags = m->list_flags;
    if (!m->half && m->variant == LJ_VARIANT_CLUSTER)
      la.flags |= LJ_LIST_CLUSTERS;
    // the cell-tile mirror serves FP64 steps on lists the library builds itself; below a few
    // hundred thousand particles the per-row kernels are faster and the mirror is not built
    if (!m->half && own_list && m->precision == LJ_PREC_FP64 && m->layout != LJ_AOS_F4 &&
        (m->variant == LJ_VARIANT_CELLTILE || (m->variant == LJ_VARIANT_AUTO && pn >= 300000)))
      la.flags |= LJ_LIST_TILES;
    if (own_list) {
      // With the tile engine the list is allocated by the build itself right after its count pass
      // (LJ_LIST_ALLOC_INTERNAL).  Otherwise the capacity protocol: a sizing call (count pass only:
      // capacity 0 returns the total), then the real build.
      la.sorted_list = nullptr; la.capacity = 0;
      const int user_flags = la.flags;
      if ((la.flags & LJ_LIST_TILES) && !(la.flags & LJ_LIST_SORT_ROWS)) la.flags |= LJ_LIST_ALLOC_INTERNAL;
      rc = lj_build_list(ctx, &la, &npairs, st);
      la.flags = user_flags;
      if (rc == LJ_OK && ctx->alloc_list) {
        list = ctx->alloc_list; capacity = ctx->alloc_capacity;
        ctx->alloc_list = nullptr;
        // the mirror was registered for the internally allocated array: same pointer from here on
        la.sorted_list = list; la.capacity = capacity;
      } else {
        if (rc != LJ_OK && rc != LJ_ERR_CAPACITY) goto done;
        capacity = npairs + npairs / 64 + 1024;  // headroom for later rebuilds
        M_TRY(lj_dev_alloc(ctx, sizeof(int32_t) * (size_t)capacity, (void**)&list, st));
        la.sorted_list = list; la.capacity = capacity;
        M_TRY(lj_build_list(ctx, &la, &npairs, st));
      }
      M_TRY(lj_list_result(ctx, &npairs, &max_np, st));
      m->list_builds = 1;
    } else {
      npairs = m->number_of_pairs_in;
      capacity = npairs;
      M_TRY(lj_dev_alloc(ctx, sizeof(int32_t) * (size_t)(capacity ? capacity : 1), (void**)&list, st));
      M_TRY(lj_upload(ctx, list, m->list_host, sizeof(int32_t) * (size_t)npairs, st));
      M_TRY(lj_upload(ctx, nop, m->number_of_partners_host, sizeof(int32_t) * (size_t)pn, st));
      M_TRY(lj_upload(ctx, ptr, m->pointer_host, sizeof(int32_t) * (size_t)pn, st));
      m->h2d_bytes += 4 * (npairs + 2 * pn);
      M_TRY(lj_validate_list(ctx, list, nop, ptr, 0, pn, npairs, st));
      // the reference's flow (cuda/force_cuda.cu:392-397): the kernel always runs on a host-built or cached
      // list.  Give it the cell-tile mirror too where that kernel wins; a system the mirror cannot take
      // (too dense for shared memory) simply stays on the per-row kernels.
      if (!m->half && m->precision == LJ_PREC_FP64 &&
          (m->variant == LJ_VARIANT_CELLTILE || (m->variant == LJ_VARIANT_AUTO && pn >= 300000))) {
        int64_t outside = 0;
        const int mrc = lj_list_mirror(ctx, q, pn, m->layout, m->plane_stride, m->search_len, nop, ptr, 0, list,
                                       capacity, 0, &outside, st);
        if (mrc != LJ_OK && m->variant == LJ_VARIANT_CELLTILE) { rc = mrc; goto done; }
      }
    }

    lj_force_args fa{};
    fa.q = q; fa.p = p; fa.pn = pn; fa.dt = m->dt; fa.cl2 = m->cl2; fa.list = list;
    fa.number_of_partners = nop; fa.pointer = ptr; fa.layout = m->layout; fa.list_layout = LJ_LIST_CSR;
    fa.variant = m->half ? LJ_VARIANT_NEWTON3 : m->variant; fa.group = m->group;
    fa.precision = m->precision; fa.pointer64 = ptr64; fa.threads_per_block = m->threads_per_block;
    fa.plane_stride = m->plane_stride;
    fa.mirror_token = lj_list_mirror_token(ctx);  // the list above is the library's own build

    if (p_on_copy_stream && cudaStreamWaitEvent(st, ctx->ev_copy, 0) != cudaSuccess) { rc = cudaErrorUnknown; goto cuda_fail; }
    M_TRY(lj_sync(ctx, st));
    const double t_k0 = now_s();
    int done_steps = 0;
    while (done_steps < m->loop) {
      int chunk = m->loop - done_steps;
      if (own_list && m->rebuild_every > 0) {
        if (done_steps > 0) {  // the list for steps [0, rebuild_every) was built above
          M_TRY(lj_build_list(ctx, &la, nullptr, st));
          fa.mirror_token = lj_list_mirror_token(ctx);
          m->list_builds++;
        }
        if (chunk > m->rebuild_every) chunk = m->rebuild_every;
      }
      M_TRY(lj_force_loop(ctx, &fa, chunk, m->use_graph, st));
      done_steps += chunk;
    }
    M_TRY(lj_sync(ctx, st));
    m->seconds_kernel = now_s() - t_k0;
    if (own_list && m->list_builds > 1) M_TRY(lj_list_result(ctx, &npairs, &max_np, st));
  }
  M_TRY(lj_download(ctx, m->p_host, p, qbytes, st));
  M_TRY(lj_sync(ctx, st));
  m->d2h_bytes += (int64_t)qbytes;
  m->number_of_pairs = npairs;
  m->max_partners = max_np;
  goto done;
cuda_fail:
  rc = lj_set_error(ctx, LJ_ERR_CUDA, "lj_measure", cudaGetErrorString((cudaError_t)rc));
done:
  if (p_on_copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  lj_dev_free(ctx, q, st);
  lj_dev_free(ctx, p, st);
  lj_dev_free(ctx, list, st);
  lj_dev_free(ctx, nop, st);
  lj_dev_free(ctx, ptr, st);
  cudaStreamSynchronize(st);
  m->seconds_total = now_s() - t_all0;
#undef M_TRY
  return rc;
}

// ------------------------------------------------------------------ multi-GPU helpers --
int lj_ipc_alloc(lj_ctx* ctx, size_t bytes, void** out) {
  LJ_ENTER(ctx);
  if (!ctx || !out) return LJ_ERR_BAD_ARG;
  *out = nullptr;
  LJ_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 16));
  return LJ_OK;
}

int lj_ipc_free(lj_ctx* ctx, void* ptr) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (ptr) LJ_CUDA(ctx, cudaFree(ptr));
  return LJ_OK;
}

int lj_ipc_export(lj_ctx* ctx, void* dev_ptr, uint8_t handle_out[64]) {
  LJ_ENTER(ctx);
  if (!ctx || !dev_ptr || !handle_out) return LJ_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  LJ_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out, &h, 64);
  return LJ_OK;
}

int lj_ipc_open(lj_ctx* ctx, const uint8_t handle[64], void** peer_ptr_out) {
  LJ_ENTER(ctx);
  if (!ctx || !handle || !peer_ptr_out) return LJ_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  LJ_CUDA(ctx, cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return LJ_OK;
}

int lj_ipc_close(lj_ctx* ctx, void* peer_ptr) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (peer_ptr) LJ_CUDA(ctx, cudaIpcCloseMemHandle(peer_ptr));
  return LJ_OK;
}

// ---- cross-GPU ordering without host round trips: 32-bit counters in peer-visible memory -------
// A rank publishes "my q of step k is final" by storing k into a flag its neighbours have mapped
// (lj_flag_set), a neighbour's pull waits for it on the device (k_copy16_sync), and reports "I have
// read your q of step k" by storing k into a flag in the OWNER's memory, which the owner waits for
// (lj_flag_wait) before it overwrites q.  System-scope release/acquire; each wait only depends on
// work of an earlier step of the other rank, so the two streams cannot wait for each other.
__global__ void k_flag_set(int* flag, int value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__device__ __forceinline__ void flag_spin(const int* flag, int at_least) {
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v < at_least) __nanosleep(200);
  } while (v < at_least);
}
__global__ void k_flag_wait(const int* flag, int at_least) { flag_spin(flag, at_least); }
// up to two segments (ghosts from the slab below and from the slab above) in ONE launch: the blocks
// are split between them, each segment waits on and reports through its own flags
struct copy_segs { lj_halo_seg s[2]; int n; };
__global__ void __launch_bounds__(256)
k_copy16_sync(const copy_segs segs, unsigned int* blocks_done) {
  const int which = (segs.n == 2 && blockIdx.x >= gridDim.x / 2) ? 1 : 0;
  const unsigned first = which ? gridDim.x / 2 : 0u;
  const unsigned nblk = segs.n == 2 ? (which ? gridDim.x - gridDim.x / 2 : gridDim.x / 2) : gridDim.x;
  const lj_halo_seg sg = segs.s[which];
  if (sg.wait_flag) {
    if (threadIdx.x == 0) flag_spin(sg.wait_flag, sg.wait_value);
    __syncthreads();
  }
  int4* __restrict__ dst = reinterpret_cast<int4*>(sg.local_dst);
  const int4* __restrict__ src = reinterpret_cast<const int4*>(sg.peer_src);
  const size_t n16 = sg.bytes / 16;
  for (size_t i = (size_t)(blockIdx.x - first) * blockDim.x + threadIdx.x; i < n16; i += (size_t)nblk * blockDim.x) {
    int4 v;  // straight from the owner's memory: no stale copy in this GPU's L1
    asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
    dst[i] = v;
  }
  if (sg.done_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(blocks_done + which, 1u) == nblk - 1) {  // the last block reports for the whole segment
        blocks_done[which] = 0;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(sg.done_flag), "r"(sg.done_value) : "memory");
      }
    }
  }
}

int lj_flag_set(lj_ctx* ctx, int32_t* flag, int32_t value, void* stream) {
  LJ_ENTER(ctx);
  LJ_REQUIRE(ctx, flag != nullptr, "lj_flag_set: null flag");
  k_flag_set<<<1, 1, 0, lj_stream(ctx, stream)>>>(flag, value);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

int lj_flag_wait(lj_ctx* ctx, const int32_t* flag, int32_t at_least, void* stream) {
  LJ_ENTER(ctx);
  LJ_REQUIRE(ctx, flag != nullptr, "lj_flag_wait: null flag");
  k_flag_wait<<<1, 1, 0, lj_stream(ctx, stream)>>>(flag, at_least);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

int lj_halo_pull_sync(lj_ctx* ctx, const lj_halo_seg* segs, int32_t nsegs, void* stream) {
  LJ_ENTER(ctx);
  LJ_REQUIRE(ctx, segs != nullptr && nsegs >= 1 && nsegs <= 2, "lj_halo_pull_sync: one or two segments");
  copy_segs cs{};
  cs.n = nsegs;
  size_t most = 0;
  for (int k = 0; k < nsegs; k++) {
    cs.s[k] = segs[k];
    LJ_REQUIRE(ctx, segs[k].bytes == 0 || (segs[k].local_dst && segs[k].peer_src), "lj_halo_pull_sync: null pointer");
    LJ_REQUIRE(ctx, segs[k].bytes % 16 == 0 && (uintptr_t)segs[k].local_dst % 16 == 0 && (uintptr_t)segs[k].peer_src % 16 == 0,
               "lj_halo_pull_sync: pointers and sizes must be multiples of 16 bytes");
    if (segs[k].bytes > most) most = segs[k].bytes;
  }
  cudaStream_t st = lj_stream(ctx, stream);
  if (!ctx->pull_counter) {
    LJ_CUDA(ctx, cudaMalloc((void**)&ctx->pull_counter, 2 * sizeof(unsigned int)));
    LJ_CUDA(ctx, cudaMemset(ctx->pull_counter, 0, 2 * sizeof(unsigned int)));
  }
  // a CTA per SM is enough to saturate an NVLink direction and leaves room for the interior force
  // kernel running concurrently; two segments share the grid
  size_t blocks = (most / 16 + 255) / 256 * (size_t)nsegs;
  if (blocks > (size_t)ctx->sm_count) blocks = (size_t)ctx->sm_count;
  if (blocks < (size_t)nsegs) blocks = (size_t)nsegs;
  k_copy16_sync<<<(unsigned)blocks, 256, 0, st>>>(cs, ctx->pull_counter);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

int lj_halo_pull(lj_ctx* ctx, void* local_dst, const void* peer_src, size_t bytes, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (bytes == 0) return LJ_OK;
  LJ_REQUIRE(ctx, local_dst && peer_src, "lj_halo_pull: null pointer");
  LJ_REQUIRE(ctx, bytes % 16 == 0 && (uintptr_t)local_dst % 16 == 0 && (uintptr_t)peer_src % 16 == 0,
             "lj_halo_pull: pointers and size must be multiples of 16 bytes");
  const size_t n16 = bytes / 16;
  size_t blocks = (n16 + 255) / 256;
  // a few CTAs per SM are enough to saturate one NVLink direction and leave the SMs to the
  // interior force kernel running concurrently
  if (blocks > (size_t)ctx->sm_count) blocks = (size_t)ctx->sm_count;
  k_copy16<<<(unsigned)blocks, 256, 0, lj_stream(ctx, stream)>>>((int4*)local_dst, (const int4*)peer_src, n16);
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

}  // extern "C"
