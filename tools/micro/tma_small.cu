// Throughput of small cp.async.bulk (TMA) copies global -> shared on every SM at once:
// one elected lane issues K copies of S bytes per mbarrier phase, R phases, double-buffered.
// Also: the same bytes moved by one warp of 16-byte cp.async (LDGSTS).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_small tma_small.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(n), "r"(smem_u32(b)) : "memory");
}

// mode 0: TMA, one lane issues all K copies; mode 1: TMA, K lanes issue one copy each (per-lane loop);
// mode 2: LDGSTS, the warp moves K*S bytes in 16-byte pieces
__global__ void k(const unsigned char* src, size_t stride, int K, int S, int R, int mode, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[2];
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const unsigned char* base = src + (size_t)blockIdx.x * stride;
  const long long t0 = clock64();
  int ph[2] = {0, 0};
  for (int r = 0; r < R + 1; r++) {
    const int b = r & 1;
    if (r < R) {
      unsigned char* dst = sm + (size_t)b * K * S;
      const unsigned char* s = base + ((size_t)r * K * S) % (stride - (size_t)K * S * 4);
      if (mode == 0) {
        if (lane == 0) { mbar_expect(&bar[b], K * S); for (int c = 0; c < K; c++) bulk(dst + c * S, s + (size_t)c * S * 3, S, &bar[b]); }
      } else if (mode == 1) {
        if (lane == 0) mbar_expect(&bar[b], K * S);
        __syncwarp();
        if (lane < K) bulk(dst + lane * S, s + (size_t)lane * S * 3, S, &bar[b]);
      } else {
        for (int o = lane * 16; o < K * S; o += 512) {
          const int c = o / S;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + o)), "l"(s + (size_t)c * S * 3 + (o - c * S)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    if (r >= 1) {
      const int pb = (r - 1) & 1;
      if (mode == 2) { if (r < R) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory"); }
      else { mbar_wait(&bar[pb], ph[pb]); ph[pb] ^= 1; }
    }
    __syncwarp();
  }
  if (lane == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const size_t stride = 8u << 20;
  unsigned char* src; cudaMalloc(&src, stride * sms); cudaMemset(src, 1, stride * sms);
  long long* out; cudaMalloc(&out, 8 * sms);
  long long* h = new long long[sms];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 3; mode++)
    for (int S : {256, 512, 1024, 2048, 4096, 8192}) {
      const int K = mode == 1 ? 10 : 10, R = 200;
      if ((size_t)2 * K * S > 200 * 1024) continue;
      k<<<sms, 32, 2 * K * S>>>(src, stride, K, S, R, mode, out);
      cudaDeviceSynchronize();
      k<<<sms, 32, 2 * K * S>>>(src, stride, K, S, R, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, 8 * sms, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
      printf("mode %d (%s) S=%5d B x K=%d per phase: %.0f cycles per phase, %.0f per copy, %.1f B/cycle/SM  %s\n", mode,
             mode == 0 ? "TMA one lane" : mode == 1 ? "TMA lane per copy" : "LDGSTS warp", S, K, avg / R, avg / R / K,
             (double)K * S * R / avg, cudaGetErrorString(e));
    }
  return 0;
}
