#!/bin/bash
mkdir -p gpurun_out
ncu --clock-control none --set full --import-source on -k regex:lj_gather_cluster_mixed -s 1 -c 1 -f -o gpurun_out/prof_cluster_mixed python tools/prof_target.py --variant cluster --prec mixed > gpurun_out/p1.log 2>&1; echo rc=$?
