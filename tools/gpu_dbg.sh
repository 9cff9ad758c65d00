#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile" 2>&1 | tail -2
for c in 16 24; do
  LJ_TILE_CONSUMERS=$c timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "force:|max\|dp|rror" | sed "s/^/[cons=$c] /"
done
for r in 32 48; do
  LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "force:|rror" | sed "s/^/[rows=$r] /"
  LJ_TILE_CONSUMERS=24 LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "force:|rror" | sed "s/^/[rows=$r cons=24] /"
done
