#!/bin/bash
for r in 40 56; do
for seg in 0 16 32; do
  LJ_TILE_SEG=$seg LJ_TILE_ROWS=$r timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed:|force:|rror" | sed "s/^/[rows=$r seg=$seg] /" | sed 's/subwarp g8 [0-9.]* ms, //; s/; build.*//; s/, per-row.*//'
done
done
