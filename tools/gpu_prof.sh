#!/bin/bash
# ncu captures (one GPU): launch list of a bench run + full-set captures of the hot kernels.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== launch list"; $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?
echo "== full: force subwarp g8"; $NCU --set full --import-source on -k regex:lj_gather_csr -s 1 -c 1 -f -o gpurun_out/prof_gather_g8 python tools/prof_target.py --variant subwarp --group 8 > gpurun_out/p1.log 2>&1; echo rc=$?
echo "== full: force tile g8"; $NCU --set full --import-source on -k regex:lj_gather_tile -s 1 -c 1 -f -o gpurun_out/prof_tile_g8 python tools/prof_target.py --variant tile --group 8 > gpurun_out/p2.log 2>&1; echo rc=$?
echo "== full: mixed g4"; $NCU --set full --import-source on -k regex:lj_gather_mixed -s 1 -c 1 -f -o gpurun_out/prof_mixed_g4 python tools/prof_target.py --variant subwarp --group 4 --prec mixed > gpurun_out/p3.log 2>&1; echo rc=$?
echo "== full: search"; $NCU --set full --import-source on -k regex:k_search -s 2 -c 2 -f -o gpurun_out/prof_search python tools/prof_target.py --steps 0 --rebuild 1 > gpurun_out/p4.log 2>&1; echo rc=$?
ls -la gpurun_out
