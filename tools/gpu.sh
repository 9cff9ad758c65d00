#!/bin/bash
# One entry point for everything that runs on a GPU box (use as: gpurun -- 'tools/gpu.sh <task> [args]').
#   check        smoke + full GPU test suite + a short bench line
#   tests [K]    pytest -m gpu (optionally -k K)
#   multi        2-GPU decomposition tests (NCCL / CUDA-IPC / C-ABI one-process) + log for profiles/
#   scale N...   bench.py --gpus N for each N given (N > 1 under torchrun), one JSON per N
#   bench [args] bench.py with clocks sampled beside it
#   ref          bench.py --impl reference (the reference's CPU path on this box's host)
#   refgpu       the reference's CUDA kernels recompiled for sm_100 (baseline/_ref) on this GPU
#   prof         round evidence: ncu launch list of the bench command + --set full captures of the
#                dominant kernels (cell-tile FP64 / mixed, k_tile_count, k_tile_replay)
#   sweep [args] tools/ct_sweep.py (needs a `make DIAG=1` build)
#   sanitize     compute-sanitizer memcheck + racecheck over every kernel family on a small system
#   micro        the micro-benchmarks under tools/micro (FP64 co-issue, small TMA copies)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
task=${1:-check}; shift || true
case "$task" in
  check)
    timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
    timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_full.log
    timeout -s KILL 600 python bench.py --steps 100 --warmup 20 --no-cpu --no-ref-gpu --no-config5 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_last.json ;;
  tests)
    timeout -s KILL 1500 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -15 ;;
  multi)
    timeout -s KILL 1200 python -m pytest tests/test_gpu_decomp.py -q 2>&1 | tail -6 | tee gpurun_out/pytest_decomp_2gpu.log ;;
  scale)
    for n in "$@"; do
      if [ "$n" -eq 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py"; fi
      timeout -s KILL 1500 $CMD --gpus "$n" --steps "${STEPS:-100}" --warmup "${WARMUP:-10}" > "gpurun_out/scale_n$n.json" 2> "gpurun_out/scale_n$n.err"; echo "bench n=$n rc=$?"
      python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/scale_n%s.json" % n) if l.startswith("{")][-1])
    print("n=%d value=%.4g ms/step=%.4f scaling=%s halo=%s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["scaling"], json.dumps(d.get("halo"))[:300]))
    w = d.get("weak_scaling_1M_per_gpu")
    if w: print("   weak 1M/GPU:", json.dumps(w)[:400])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/scale_n%s.err" % n).read()[-1500:])
PY
    done ;;
  bench)
    nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv &
    SMI=$!
    timeout -s KILL 1500 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
    kill $SMI; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
  ref)
    timeout -s KILL 900 python bench.py --impl reference --steps "${STEPS:-40}" --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json ;;
  refgpu)
    timeout -s KILL 900 python baseline/run_reference_gpu.py --json gpurun_out/reference_gpu.json | tail -40 ;;
  prof)
    $NCU --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 200 --warmup 20 --no-cpu --no-ref-gpu --no-config5 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
    $NCU --set full --import-source on -k regex:lj_celltile_force -s 2 -c 1 -f -o gpurun_out/prof_celltile python tools/prof_target.py --variant auto --steps 4 > gpurun_out/p1.log 2>&1; echo "full fp64 rc=$?"
    $NCU --set full --import-source on -k regex:lj_celltile_force -s 2 -c 1 -f -o gpurun_out/prof_celltile_mixed python tools/prof_target.py --variant auto --prec mixed --wide --steps 4 > gpurun_out/p2.log 2>&1; echo "full mixed rc=$?"
    $NCU --set full --import-source on -k "regex:k_tile_count|k_tile_replay" -s 2 -c 2 -f -o gpurun_out/prof_tile_engine python tools/prof_target.py --variant auto --steps 0 --rebuild 2 > gpurun_out/p3.log 2>&1; echo "full build rc=$?"
    $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/build_launches.csv python tools/prof_target.py --variant auto --steps 0 --rebuild 2 > gpurun_out/p4.log 2>&1; echo "build launch list rc=$?" ;;
  sweep)
    nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
    python tools/ct_sweep.py "$@" 2>&1 | tee gpurun_out/ct_sweep.log ;;
  sanitize)
    export PATH=/usr/local/cuda/bin:$PATH
    timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/debug/sanitize_target.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
    # racecheck does not model the mbarrier / TMA pipeline of lj_celltile_force (it reports the producer's bulk
    # copies against consumer reads that the tile barriers order; profiles/README.md): every other kernel here,
    # the cell-tile kernel is covered by memcheck above and by the bit-exact tests
    timeout -s KILL 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/debug/sanitize_target.py --skip-celltile > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log ;;
  micro)
    for m in dp_coissue tma_small fp64_peak dp_latency; do [ -x tools/micro/$m ] && ./tools/micro/$m | tee gpurun_out/micro_$m.txt; done ;;
  *) echo "unknown task $task"; exit 2 ;;
esac
