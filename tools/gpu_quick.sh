#!/bin/bash
# quick: subset of tests + build timing + a few force variants
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "${TESTK:-list or cluster or full_size or measure or config_B}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_quick.log
timeout -s KILL 600 python tools/sweep.py --quick --reps 10 > gpurun_out/sweep_quick.log 2>&1; echo "sweep rc=$?"; head -2 gpurun_out/sweep_quick.log; grep -E "cluster|subwarp  g=8  tb=128|mixed" gpurun_out/sweep_quick.log
