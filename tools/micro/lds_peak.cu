// Shared-memory load microbenchmark: cycles per warp-level LDS for several widths and index
// patterns (what does a 24-byte-record gather really cost on the B200 data pipe?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_peak lds_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// pattern 0: lane-consecutive records; 1: 8-lane runs at random bases (4 groups); 2: fully random
template <int WIDTH>  // 64: three LDS.64 per record (24 B stride); 128: LDS.128 + LDS.64 on 32 B records
__global__ void k_lds(double* out, int iters, int pattern, int nrec) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < nrec * 4; i += blockDim.x) sm[i] = i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t state = threadIdx.x * 2654435761u + 12345u;
  const uint32_t base = smem_u32(sm);
  double ax = 0, ay = 0, az = 0;
  for (int it = 0; it < iters; it++) {
    uint32_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      state = state * 1664525u + 1013904223u;
      uint32_t r = state >> 8;
      if (pattern == 0) idx[u] = (it * 4 + u) * 32 + lane;
      else if (pattern == 1) idx[u] = __shfl_sync(0xffffffffu, r, lane & ~7) + (lane & 7);
      else idx[u] = r;
      idx[u] %= (uint32_t)(nrec - 8);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      double x, y, z;
      if (WIDTH == 64) {
        const uint32_t a = base + idx[u] * 24u;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(y) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(z) : "r"(a));
      } else {
        const uint32_t a = base + idx[u] * 32u;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(z) : "r"(a));
      }
      ax += x; ay += y; az += z;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ax + ay + az;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  const int nrec = 4096, iters = 4000;
  const size_t smem = (size_t)nrec * 32;
  cudaFuncSetAttribute(k_lds<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_lds<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int width : {64, 128})
    for (int pattern : {0, 1, 2})
      for (int warps : {8, 16, 32}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&] {
          if (width == 64) k_lds<64><<<sms, warps * 32, smem>>>(out, iters, pattern, nrec);
          else k_lds<128><<<sms, warps * 32, smem>>>(out, iters, pattern, nrec);
        };
        run(); cudaDeviceSynchronize();
        cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double recs = (double)warps * iters * 4;  // warp-level record fetches per SM
        printf("width %3d pattern %d warps %2d: %.3f ms -> %.2f cycles per warp-record (3 doubles x 32 lanes)\n",
               width, pattern, warps, ms, ms * 1e-3 * 1.965e9 / recs);
      }
  return 0;
}
