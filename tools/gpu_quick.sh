#!/bin/bash
# quick: list tests + build timing + a few force variants
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "list or smoke or full_size or measure" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_quick.log
timeout -s KILL 600 python tools/sweep.py --quick --reps 10 > gpurun_out/sweep_quick.log 2>&1; echo "sweep rc=$?"; head -3 gpurun_out/sweep_quick.log; grep -E "g=8 |g=4 " gpurun_out/sweep_quick.log
