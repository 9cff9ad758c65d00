#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:lj_gather_cluster -s 1 -c 1 -f -o gpurun_out/prof_cluster_b python tools/prof_target.py --variant cluster --group 32 > gpurun_out/p1.log 2>&1; echo rc=$?
$NCU --set full --import-source on -k regex:lj_gather_cluster_lanes -s 1 -c 1 -f -o gpurun_out/prof_cluster_c python tools/prof_target.py --variant cluster > gpurun_out/p2.log 2>&1; echo rc=$?
