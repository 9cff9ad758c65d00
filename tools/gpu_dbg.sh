#!/bin/bash
for mb in 2 3 4; do
LJ_TFC_MB=$mb timeout -s KILL 300 python tools/celltile_check.py --reps 10 2>&1 | grep -E "force:|rror" | sed "s/^/[mb=$mb] /" | sed 's/force: subwarp.*build/build/'
done
