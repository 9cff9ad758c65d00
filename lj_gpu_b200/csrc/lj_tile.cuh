// lj_tile.cuh -- mbarrier / TMA bulk-copy helpers and the transposing batch reduction shared by
// the CTA-tile kernels (lj_force_tile.cu, lj_force_cluster.cu).
#pragma once
#include "lj_common.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LJ_WAIT:\n"
#ifdef LJ_MBAR_HINT_NS  // suspend-time hint: the warp sleeps in hardware until the phase completes or the time is up
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
      "@p bra LJ_DONE;\n"
      "bra LJ_WAIT;\n"
      "LJ_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
#ifdef LJ_MBAR_HINT_NS
      , "r"((uint32_t)LJ_MBAR_HINT_NS)
#endif
      : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }

// Sum over the G lanes of a group for B rows at once.  On return lane l holds in `out` the total
// of row  ((l & G/2) ? B/2 : 0) + ((l & G/4) ? B/4 : 0) ...  (the bits consumed by the
// transposing steps); every lane of the group holds a valid total for "its" row.
template <int G, int B>
__device__ __forceinline__ double batch_sum(const double (&a)[B], int lg, unsigned gmask,
                                            int& my_row) {
  static_assert(B == 1 || B == 2 || B == 4, "batch of 1, 2 or 4 rows");
  double v;
  int stride = G / 2;
  my_row = 0;
  if (B == 4) {
    const bool up = (lg & stride) != 0;
    double k0 = sel(up, a[2], a[0]) + __shfl_xor_sync(gmask, sel(up, a[0], a[2]), stride);
    double k1 = sel(up, a[3], a[1]) + __shfl_xor_sync(gmask, sel(up, a[1], a[3]), stride);
    my_row = up ? 2 : 0;
    stride >>= 1;
    const bool up2 = (lg & stride) != 0;
    v = sel(up2, k1, k0) + __shfl_xor_sync(gmask, sel(up2, k0, k1), stride);
    my_row += up2 ? 1 : 0;
    stride >>= 1;
  } else if (B == 2) {
    const bool up = (lg & stride) != 0;
    v = sel(up, a[1], a[0]) + __shfl_xor_sync(gmask, sel(up, a[0], a[1]), stride);
    my_row = up ? 1 : 0;
    stride >>= 1;
  } else {
    v = a[0];
  }
  for (; stride >= 1; stride >>= 1) v += __shfl_xor_sync(gmask, v, stride);
  return v;
}

