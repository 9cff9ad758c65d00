#!/bin/bash
# One GPU-box session: smoke, parity tests, variant sweep, bench, ncu launch list.
# Usage: tools/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/gpu.txt
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
echo "== sweep"; timeout -s KILL 600 python tools/sweep.py $SWEEP_ARGS > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep.log | tail -80
echo "== bench"; timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
