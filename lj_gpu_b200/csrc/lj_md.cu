// lj_md.cu -- the caller the hot path is meant for: an MD step around the force call.
//
// Not in the reference (its q is static, cuda/force_cuda.cu:333-335 applies the same force LOOP
// times).  The force call is the "kick" p += F dt of a symplectic Euler step; this file adds
//   lj_drift              q += p dt                       (unit mass)
//   lj_max_displacement2  max_i |q_i - q_ref,i|^2         (skin-based rebuild trigger: rebuild when
//                                                          it exceeds ((search - cutoff)/2)^2)
//   lj_energy             kinetic sum p^2/2 and potential sum 4(r^-12 - r^-6) over listed pairs
//                         within the cutoff               (conservation checks)
// All three are plain streaming / gather kernels; the potential reuses the force kernel's list walk.
#include "lj_common.cuh"

namespace {

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_drift(void* __restrict__ q, const void* __restrict__ p, int64_t pn, int64_t plane, double dt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pn) return;
  double px, py, pz;
  load_pos<LAYOUT>(p, i, plane, px, py, pz);
  add_mom<LAYOUT>(q, i, plane, px * dt, py * dt, pz * dt);  // same in-place update helper, .w kept
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_max_disp2(const void* __restrict__ q, const void* __restrict__ qref, int64_t pn, int64_t plane,
            unsigned long long* out) {
  double m = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pn;
       i += (int64_t)gridDim.x * blockDim.x) {
    double x, y, z, xr, yr, zr;
    load_pos<LAYOUT>(q, i, plane, x, y, z);
    load_pos<LAYOUT>(qref, i, plane, xr, yr, zr);
    const double dx = x - xr, dy = y - yr, dz = z - zr;
    m = fmax(m, fma(dz, dz, fma(dy, dy, dx * dx)));
  }
  for (int s = 16; s >= 1; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_kinetic(const void* __restrict__ p, int64_t pn, int64_t plane, double* out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pn;
       i += (int64_t)gridDim.x * blockDim.x) {
    double x, y, z;
    load_pos<LAYOUT>(p, i, plane, x, y, z);
    s += 0.5 * (x * x + y * y + z * z);
  }
  for (int k = 16; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// potential energy over the listed pairs with r2 <= cl2, 8 lanes per row (the force kernel's walk)
template <int LAYOUT, bool PTR64>
__global__ void __launch_bounds__(256)
k_potential(const void* __restrict__ q, int64_t pn, int64_t plane, double cl2,
            const int32_t* __restrict__ list, const int32_t* __restrict__ nop,
            const void* __restrict__ pointer, double* out) {
  constexpr int G = 8;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int lg = threadIdx.x % G;
  double e = 0.0;
  if (i < pn) {
    double xi, yi, zi;
    load_pos<LAYOUT>(q, i, plane, xi, yi, zi);
    const int np = __ldg(nop + i);
    const int32_t* __restrict__ row = list + row_offset<PTR64>(pointer, i);
    for (int k = lg; k < np; k += G) {
      const int j = __ldg(row + k);
      double xj, yj, zj;
      load_pos<LAYOUT>(q, j, plane, xj, yj, zj);
      const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
      const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      if (r2 <= cl2) {
        const double x = 1.0 / r2, x3 = x * x * x;
        e += 4.0 * (x3 * x3 - x3);
      }
    }
  }
  for (int k = 16; k >= 1; k >>= 1) e += __shfl_xor_sync(0xffffffffu, e, k);
  if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(out, e);
}

int blocks_of(int64_t n, int tb, int cap) {
  int64_t b = (n + tb - 1) / tb;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" {

int lj_drift(lj_ctx* ctx, void* q, const void* p, int64_t pn, int32_t layout, int64_t plane_stride,
             double dt, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  if (pn <= 0) return LJ_OK;
  LJ_REQUIRE(ctx, q && p, "lj_drift: null array");
  cudaStream_t st = lj_stream(ctx, stream);
  const unsigned blocks = (unsigned)((pn + 255) / 256);
  switch (layout) {
    case LJ_AOS_D3: k_drift<LJ_AOS_D3><<<blocks, 256, 0, st>>>(q, p, pn, plane_stride, dt); break;
    case LJ_AOS_D4: k_drift<LJ_AOS_D4><<<blocks, 256, 0, st>>>(q, p, pn, plane_stride, dt); break;
    case LJ_SOA_D: k_drift<LJ_SOA_D><<<blocks, 256, 0, st>>>(q, p, pn, plane_stride, dt); break;
    default: return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_drift", "layout");
  }
  LJ_LAUNCHED(ctx);
  return LJ_OK;
}

int lj_max_displacement2(lj_ctx* ctx, const void* q, const void* q_ref, int64_t pn, int32_t layout,
                         int64_t plane_stride, double* out_host, void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, out_host != nullptr, "lj_max_displacement2: null output");
  *out_host = 0.0;
  if (pn <= 0) return LJ_OK;
  LJ_REQUIRE(ctx, q && q_ref, "lj_max_displacement2: null array");
  cudaStream_t st = lj_stream(ctx, stream);
  int rc = lj_scratch_reserve(ctx, 1, st);
  if (rc) return rc;
  unsigned long long* slot = reinterpret_cast<unsigned long long*>(ctx->bbox) + 6;  // spare word
  LJ_CUDA(ctx, cudaMemsetAsync(slot, 0, 8, st));
  const int blocks = blocks_of(pn, 256, 8 * ctx->sm_count);
  switch (layout) {
    case LJ_AOS_D3: k_max_disp2<LJ_AOS_D3><<<blocks, 256, 0, st>>>(q, q_ref, pn, plane_stride, slot); break;
    case LJ_AOS_D4: k_max_disp2<LJ_AOS_D4><<<blocks, 256, 0, st>>>(q, q_ref, pn, plane_stride, slot); break;
    case LJ_SOA_D: k_max_disp2<LJ_SOA_D><<<blocks, 256, 0, st>>>(q, q_ref, pn, plane_stride, slot); break;
    default: return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_max_displacement2", "layout");
  }
  LJ_LAUNCHED(ctx);
  LJ_CUDA(ctx, cudaMemcpyAsync(out_host, slot, 8, cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  return LJ_OK;
}

int lj_energy(lj_ctx* ctx, const lj_force_args* a, double* kinetic_out, double* potential_out,
              void* stream) {
  LJ_ENTER(ctx);
  if (!ctx) return LJ_ERR_BAD_ARG;
  LJ_REQUIRE(ctx, a && kinetic_out && potential_out, "lj_energy: null argument");
  *kinetic_out = *potential_out = 0.0;
  if (a->pn <= 0) return LJ_OK;
  LJ_REQUIRE(ctx, a->q && a->p && a->list && a->number_of_partners && a->pointer &&
                      a->list_layout == LJ_LIST_CSR,
             "lj_energy: needs q, p and a CSR list");
  cudaStream_t st = lj_stream(ctx, stream);
  int rc = lj_scratch_reserve(ctx, 1, st);
  if (rc) return rc;
  double* acc = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->grid) + 192);  // two spare doubles
  LJ_CUDA(ctx, cudaMemsetAsync(acc, 0, 16, st));
  const int rb = blocks_of(a->pn, 256, 8 * ctx->sm_count);
  const unsigned pb = (unsigned)((a->pn * 8 + 255) / 256);
#define LJ_E_CASE(LAY)                                                                              \
  k_kinetic<LAY><<<rb, 256, 0, st>>>(a->p, a->pn, a->plane_stride, acc);                             \
  if (a->pointer64)                                                                                  \
    k_potential<LAY, true><<<pb, 256, 0, st>>>(a->q, a->pn, a->plane_stride, a->cl2, a->list,        \
                                                a->number_of_partners, a->pointer, acc + 1);         \
  else                                                                                               \
    k_potential<LAY, false><<<pb, 256, 0, st>>>(a->q, a->pn, a->plane_stride, a->cl2, a->list,       \
                                                 a->number_of_partners, a->pointer, acc + 1);
  switch (a->layout) {
    case LJ_AOS_D3: LJ_E_CASE(LJ_AOS_D3) break;
    case LJ_AOS_D4: LJ_E_CASE(LJ_AOS_D4) break;
    case LJ_SOA_D: LJ_E_CASE(LJ_SOA_D) break;
    default: return lj_set_error(ctx, LJ_ERR_BAD_ARG, "lj_energy", "layout");
  }
#undef LJ_E_CASE
  ctx->launches += 2;
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return lj_set_error(ctx, LJ_ERR_CUDA, "lj_energy launch", cudaGetErrorString(e));
  }
  double host[2];
  LJ_CUDA(ctx, cudaMemcpyAsync(host, acc, 16, cudaMemcpyDeviceToHost, st));
  LJ_CUDA(ctx, cudaStreamSynchronize(st));
  *kinetic_out = host[0];
  // a full (directed) list visits every pair twice, a half list (variant NEWTON3) once
  *potential_out = a->variant == LJ_VARIANT_NEWTON3 ? host[1] : 0.5 * host[1];
  return LJ_OK;
}

}  // extern "C"
