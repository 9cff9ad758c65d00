#!/bin/bash
# cell-tile kernel sweep on the GPU box (DIAG build expected in tree)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python tools/ct_sweep.py "$@" 2>&1 | tee gpurun_out/ct_sweep.log
