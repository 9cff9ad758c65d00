#!/bin/bash
# mixed-precision cell-tile kernel: parity tests, timing with consumer-warp / unroll variants, one
# full ncu capture
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile" > gpurun_out/pytest_mx.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_mx.log
for c in 16 24 31; do
  LJ_TILE_CONSUMERS=$c timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed|force:" | sed "s/^/[cons=$c] /"
done
for v in u8 u2; do
  LJ_B200_LIB=build_variants/liblj_b200_$v.so timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed" | sed "s/^/[$v] /"
done
for r in 24 64; do
  LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed|force:" | sed "s/^/[rows=$r] /"
done
LJ_TILE_DBG=1 timeout -s KILL 300 python tools/prof_target.py --variant celltile --prec mixed --steps 6 2>&1 | tail -3
timeout -s KILL 600 ncu --clock-control none --set full --import-source on -k regex:lj_celltile_force -s 2 -c 1 -f -o gpurun_out/prof_mx_celltile python tools/prof_target.py --variant celltile --prec mixed --steps 4 > gpurun_out/p_mx.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 600 python bench.py --steps 100 --warmup 20 --no-cpu > gpurun_out/bench_mx1.json 2> gpurun_out/bench_mx1.err; echo "bench rc=$?"; cat gpurun_out/bench_mx1.json
