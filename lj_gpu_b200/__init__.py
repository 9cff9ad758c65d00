"""lj_gpu_b200 -- B200-native Lennard-Jones force + neighbour-list hot path of kohnakagawa/lj_gpu.

The product is lj_gpu_b200/liblj_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/lj_b200.h) plus the C++ host driver lj_gpu_b200/driver/force_b200.  This package is
the thin ctypes mirror used by tests and bench.py.  No CPU fallback exists.
"""
from . import _capi  # noqa: F401
from .api import (CL2, CUTOFF_LENGTH, DENSITY, DT, LOOP, L_BOX, SEARCH_LENGTH, CudaPtr, LJContext,
                  LJError, PairList, init_fcc, loadpair, loadpair_dat, makepaircache, print_results,
                  savepair_dat)

__all__ = ["LJContext", "LJError", "PairList", "CudaPtr", "init_fcc", "print_results", "makepaircache",
           "loadpair", "savepair_dat", "loadpair_dat", "DENSITY",
           "L_BOX", "DT", "CUTOFF_LENGTH", "SEARCH_LENGTH", "CL2", "LOOP"]
