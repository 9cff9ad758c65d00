// FP64 pipe microbenchmark: DFMA throughput per SM for a given number of warps and ILP,
// and the lj_pair body itself (17 FP64 ops + MUFU + compare) on register data.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lj_gpu_b200/csrc/lj_common.cuh"

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_pair(double* out, int iters, double c24, double c48, long long cl2) {
  double x[ILP], fx = 0, fy = 0, fz = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = 1.0 + threadIdx.x * 1e-3 + i * 0.1;
  const double xi = 0.25, yi = 0.5, zi = 0.75;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      lj_pair(x[i] - xi, x[i] * 0.5 - yi, x[i] * 0.25 - zi, c24, c48, cl2, fx, fy, fz);
      x[i] += 1e-9;
    }
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
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 2048);
  const int iters = 20000;
  double c = 9.0; long long cl2; memcpy(&cl2, &c, 8);
  printf("SMs %d, clock attr %d kHz\n", sms, clk_khz);
  for (int warps : {4, 8, 16, 32, 64}) {
    const int tb = warps * 32 > 1024 ? 1024 : warps * 32;
    const int blocks = sms * (warps * 32 / tb);
    float ms = timeit([&] { k_dfma<8><<<blocks, tb>>>(out, iters, 1.0000001, 1e-9); });
    double ops = (double)blocks * tb * iters * 8;
    printf("DFMA ILP8  warps/SM %2d: %.3f ms  -> %.1f DFMA lanes/clk/SM (at 1.965 GHz)\n", warps, ms,
           ops / (ms * 1e-3) / sms / 1.965e9);
    ms = timeit([&] { k_pair<4><<<blocks, tb>>>(out, iters / 10, 24e-3, 48e-3, cl2); });
    double pairs = (double)blocks * tb * (iters / 10) * 4;
    printf("lj_pair ILP4 warps/SM %2d: %.3f ms  -> %.2f cycles per warp-pair-instruction group (32 pairs) per SM, %.3e pairs/s\n",
           warps, ms, (ms * 1e-3 * 1.965e9) / (pairs / 32 / sms), pairs / (ms * 1e-3));
  }
  return 0;
}
