// FP64 on B200: (1) dependent-issue latency of DFMA (C independent chains per warp, W warps per SM
// sub-partition); (2) the LJ pair body as the cell-tile kernel runs it (17 FP64 instructions, MUFU.RCP64H,
// bit-pattern cutoff) on register operands only -- no shared-memory loads, no list -- with U pairs per
// lane in flight and W warps per sub-partition: how close to the FP64-pipe floor (17 x 2 cycles per 32
// pairs) does the arithmetic alone get?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dp_latency dp_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int C>
__global__ void k_chain(double* out, int iters, double a, double b) {
  double v[C];
#pragma unroll
  for (int i = 0; i < C; i++) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < C; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void lj_pair(double dx, double dy, double dz, double c24, double c48, long long cl2_bits,
                                        double& fx, double& fy, double& fz) {
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(r2));
  x0 = __hiloint2double((__double_as_longlong(r2) <= cl2_bits) ? __double2hiint(x0) : 0, 0);
  const double e = fma(-r2, x0, 1.0);
  const double x = fma(x0, fma(e, e, e), x0);
  const double x3 = x * x * x;
  const double df = (x * x3) * fma(-c48, x3, c24);
  fx = fma(df, dx, fx);
  fy = fma(df, dy, fy);
  fz = fma(df, dz, fz);
}

// positions advance by a register increment per iteration: the only non-pair instructions are U x 3 DADD
// (counted: they stand in for the loads' address arithmetic, they are FP64 too, so the report subtracts them)
template <int U>
__global__ void k_pair(double* out, int iters, double c24, double c48, long long cl2_bits, double step) {
  double xj[U], yj[U], zj[U];
#pragma unroll
  for (int u = 0; u < U; u++) { xj[u] = 1.0 + 0.01 * threadIdx.x + u; yj[u] = 0.5 + u; zj[u] = 0.25 * u; }
  const double xi = 0.1, yi = 0.2, zi = 0.3;
  double fx = 0, fy = 0, fz = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < U; u++) lj_pair(xj[u] - xi, yj[u] - yi, zj[u] - zi, c24, c48, cl2_bits, fx, fy, fz);
#pragma unroll
    for (int u = 0; u < U; u++) xj[u] += step;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = fx + fy + fz;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double hz = khz * 1e3;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 2048);
  const int iters = 4000;
  printf("SM clock %.0f MHz (attribute), %d SMs\n", hz / 1e6, sms);
  auto chain = [&](auto tag, int wps) {
    constexpr int C = decltype(tag)::value;
    const int tb = wps * 4 * 32;
    const float ms = timeit([&] { k_chain<C><<<sms, tb>>>(out, iters, 1.0000001, 1e-9); });
    const double cyc = ms * 1e-3 * hz / ((double)iters * 8 * C * wps);
    printf("chain  C=%d chains/warp, %d warps/sub-partition: %.2f cycles per DFMA per sub-partition (C=1, 1 warp: the latency)\n", C, wps, cyc);
  };
  for (int wps : {1, 2, 4}) {
    chain(std::integral_constant<int, 1>{}, wps);
    chain(std::integral_constant<int, 2>{}, wps);
    chain(std::integral_constant<int, 4>{}, wps);
    chain(std::integral_constant<int, 8>{}, wps);
  }
  auto pair = [&](auto tag, int wps) {
    constexpr int U = decltype(tag)::value;
    const int tb = wps * 4 * 32;
    const float ms = timeit([&] { k_pair<U><<<sms, tb>>>(out, iters, 0.024, 0.048, 0x4022000000000000LL, 1e-4); });
    const double cyc = ms * 1e-3 * hz / ((double)iters * U * wps);
    printf("pair   U=%d pairs in flight, %d warps/sub-partition: %.1f cycles per 32 pairs per sub-partition "
           "(17 + 1 FP64 instructions: floor %.1f)\n", U, wps, cyc, 18 * 2.0);
  };
  for (int wps : {1, 2, 4, 6, 8}) {
    pair(std::integral_constant<int, 1>{}, wps);
    pair(std::integral_constant<int, 2>{}, wps);
    pair(std::integral_constant<int, 4>{}, wps);
    pair(std::integral_constant<int, 8>{}, wps);
  }
  return 0;
}
