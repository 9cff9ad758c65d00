#!/bin/bash
# round-1 evidence (final state): bench line, ncu launch list of the same command, full captures of
# the dominant kernels (dram bytes = "traffic"), clocks during the run, sanitizer on the new kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
kill $SMI
cat gpurun_out/bench_final.json
python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
$NCU --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
$NCU --set full --import-source on -k regex:lj_celltile_force -s 2 -c 1 -f -o gpurun_out/prof_final_celltile python tools/prof_target.py --variant auto --steps 4 > gpurun_out/p1.log 2>&1; echo "full rc=$?"
$NCU --set full --import-source on -k regex:lj_celltile_force -s 2 -c 1 -f -o gpurun_out/prof_final_celltile_mixed python tools/prof_target.py --variant auto --prec mixed --wide --steps 4 > gpurun_out/p2.log 2>&1; echo "full rc=$?"
$NCU --set full --import-source on -k "regex:k_search_cluster|k_tile_fill" -s 2 -c 2 -f -o gpurun_out/prof_final_search python tools/prof_target.py --variant auto --steps 0 --rebuild 1 > gpurun_out/p3.log 2>&1; echo "full rc=$?"
timeout -s KILL 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "celltile_mixed or six_array" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
