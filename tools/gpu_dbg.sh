#!/bin/bash
mkdir -p gpurun_out
for ry in 0 10; do
LJ_TILE_RY=$ry timeout -s KILL 200 python tools/debug/mx_nan.py 48 2>&1 | grep -v "^  row\|worst" | tail -4
done
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile" > gpurun_out/pytest_mx.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mx.log
for r in 40 64 96 128; do
  LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed|force:" | sed "s/^/[rows=$r] /"
done
