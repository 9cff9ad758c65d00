// Does an FP64 instruction (half-rate pipe, 16 lanes per SM sub-partition) hold the issue port for
// its two pipe cycles, or can the scheduler issue to another pipe on the alternate cycle?
// Per loop iteration: 8 independent DFMA + K independent non-DP instructions of a given kind.
// If co-issue works, time stays flat until K reaches 8 (one free slot per DFMA); if not, every
// extra instruction adds a cycle.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dp_coissue dp_coissue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void k_mix(double* out, int iters, double a, double b, int ia, float fa) {
  __shared__ double sh[1024];
  sh[threadIdx.x & 1023] = threadIdx.x;
  __syncthreads();
  double v[8];
  int w[16];
  float f[16];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 16; i++) { w[i] = threadIdx.x + i; f[i] = threadIdx.x * 0.5f + i; }
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh) + (threadIdx.x & 31) * 8;
  double ld = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(a), "d"(b));
#pragma unroll
      for (int k = 0; k < K / 8 + (i < K % 8 ? 1 : 0); k++) {
        const int s = (i * 2 + k) & 15;
        if (KIND == 0) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(w[s]) : "r"(ia));
        if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[s]) : "f"(fa));
        if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(w[s]) : "r"(ia));
        if (KIND == 3) {
          double t;
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(sbase + (uint32_t)(s * 256)));
          ld += t;  // one DADD per load: counted in the report
        }
      }
    }
  }
  double s = ld;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
#pragma unroll
  for (int i = 0; i < 16; i++) s += w[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

template <int K, int KIND>
void run(double* out, int sms, int warps, const char* kind) {
  const int iters = 20000, tb = warps * 32 > 1024 ? 1024 : warps * 32;
  const int blocks = sms * (warps * 32 / tb);
  const float ms = timeit([&] { k_mix<K, KIND><<<blocks, tb>>>(out, iters, 1.0000001, 1e-9, 3, 1.0001f); });
  // cycles per iteration per SM sub-partition (warps / 4 warps each run `iters` iterations)
  const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * warps / 4.0);
  printf("%-5s K=%2d warps/SM %2d: %.3f ms -> %.2f cycles per (8 DFMA + %d %s) per sub-partition\n", kind, K, warps,
         ms, cyc, K, kind);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 2048);
  for (int warps : {16, 32}) {
    run<0, 0>(out, sms, warps, "imad");
    run<4, 0>(out, sms, warps, "imad");
    run<8, 0>(out, sms, warps, "imad");
    run<16, 0>(out, sms, warps, "imad");
    run<8, 1>(out, sms, warps, "ffma");
    run<16, 1>(out, sms, warps, "ffma");
    run<8, 2>(out, sms, warps, "lop3");
    run<16, 2>(out, sms, warps, "lop3");
    run<4, 3>(out, sms, warps, "lds64");
    run<8, 3>(out, sms, warps, "lds64");
  }
  return 0;
}
