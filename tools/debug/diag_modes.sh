#!/bin/bash
# Kernel-surgery modes and per-warp counters of the cell-tile kernel on a DIAG variant build:
#   tools/build_variants.sh diag "-DLJ_DIAG=1" && gpurun -- bash tools/debug/diag_modes.sh
export LJ_B200_LIB=$PWD/build_variants/liblj_b200_diag.so
python tools/ct_sweep.py --reps 50 --check-steps 3 --prec fp64 --configs ";LJ_TILE_MODE=3;LJ_TILE_MODE=19;LJ_TILE_MODE=35;LJ_TILE_MODE=115;LJ_TILE_MODE=1" 2>&1 | grep -E "rows="
LJ_TILE_DBG=1 python tools/ct_sweep.py --reps 6 --check-steps 3 --prec fp64 --configs "" 2>&1 | grep -E "dbg|rows=" | head -8
