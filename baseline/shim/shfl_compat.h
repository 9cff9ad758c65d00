// Force-included (-include) ahead of the reference's cuda/force_cuda.cu: the pre-Volta
// __shfl_down(var, delta[, width]) the reference uses (cuda/device_util.cuh:13-17) no longer exists
// for sm_70+; map it to the full-mask _sync form.  Variadic, so that toolkit headers that still
// mention the old name with three arguments (cuda_fp16.hpp) keep compiling.
#pragma once
#include <cuda_runtime.h>
template <typename T>
__device__ __forceinline__ T lj_shfl_down_compat(T v, unsigned k) { return __shfl_down_sync(0xffffffffu, v, k); }
template <typename T>
__device__ __forceinline__ T lj_shfl_down_compat(T v, unsigned k, int w) { return __shfl_down_sync(0xffffffffu, v, k, w); }
#define __shfl_down(...) lj_shfl_down_compat(__VA_ARGS__)
