#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile" > gpurun_out/pytest_mx.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mx.log
LJ_B200_LIB=build_variants/liblj_b200_g4.so timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile" > gpurun_out/pytest_g4.log 2>&1; echo "pytest g4 rc=$?"; tail -3 gpurun_out/pytest_g4.log
for r in 40 56; do
  LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed|rror" | sed "s/^/[rows=$r g8] /"
for v in g4 g4u8; do
  LJ_B200_LIB=build_variants/liblj_b200_$v.so LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed|rror" | sed "s/^/[rows=$r $v] /"
done
done
