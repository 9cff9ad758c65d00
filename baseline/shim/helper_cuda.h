// Stand-in for the CUDA samples header the reference includes (cuda/cuda_ptr.cuh:8) -- the samples
// are not part of the toolkit.  Only `checkCudaErrors` is used by the reference; same behaviour as
// the original: print the error and exit(EXIT_FAILURE).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <cuda_runtime.h>
#define checkCudaErrors(call)                                                              \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      std::fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\"\n", __FILE__, __LINE__, \
                   (int)e__, cudaGetErrorName(e__), #call);                                \
      std::exit(EXIT_FAILURE);                                                             \
    }                                                                                      \
  } while (0)
