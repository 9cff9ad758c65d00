#!/bin/bash
# full GPU suite + smoke + a short bench: the last check before a commit that touches kernels
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_full.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 300 python bench.py --steps 100 --warmup 20 --no-cpu > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_last.json
