#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:lj_gather_tile -s 1 -c 1 -f -o gpurun_out/prof_tile2_g32 python tools/prof_target.py --variant tile --group 32 > gpurun_out/p1.log 2>&1; echo rc=$?
$NCU --set full --import-source on -k regex:lj_gather_tile -s 1 -c 1 -f -o gpurun_out/prof_tile2_g8 python tools/prof_target.py --variant tile --group 8 > gpurun_out/p2.log 2>&1; echo rc=$?
$NCU --set full --import-source on -k regex:k_search -s 2 -c 2 -f -o gpurun_out/prof_search2 python tools/prof_target.py --steps 0 --rebuild 1 > gpurun_out/p4.log 2>&1; echo rc=$?
$NCU --metrics gpu__time_duration.sum -k regex:k_ -s 17 -c 17 --csv --log-file gpurun_out/build_launches.csv python tools/prof_target.py --steps 0 --rebuild 1 > gpurun_out/p5.log 2>&1; echo rc=$?
