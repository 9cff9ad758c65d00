#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:lj_gather_cluster -s 1 -c 1 -f -o gpurun_out/prof_cluster python tools/prof_target.py --variant cluster > gpurun_out/p1.log 2>&1; echo rc=$?
$NCU --set full --import-source on -k regex:k_search_cluster -s 2 -c 2 -f -o gpurun_out/prof_search_cl python tools/prof_target.py --steps 0 --rebuild 1 > gpurun_out/p4.log 2>&1; echo rc=$?
