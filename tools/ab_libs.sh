#!/bin/bash
# A/B timing of build variants on one box: tools/ab_libs.sh [name ...]   (build_variants/liblj_b200_<name>.so; "prod" = the product library)
# Each library is checked bit for bit (FP64) / to 1e-5 (mixed) against the per-row kernel over 30 steps, then timed (tools/ct_sweep.py).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for name in "$@"; do
  if [ "$name" = prod ]; then lib=lj_gpu_b200/liblj_b200.so; else lib=build_variants/liblj_b200_$name.so; fi
  echo "== $name"
  LJ_B200_LIB=$PWD/$lib timeout -s KILL 300 python tools/ct_sweep.py --configs "" --reps "${REPS:-100}" ${SWEEP_ARGS:-} 2>&1 | grep -E "rows=|Error|error|Traceback" | tee -a gpurun_out/ab_libs.log
done
