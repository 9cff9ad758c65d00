#!/bin/bash
# multi-GPU session: decomposition parity test + scaling bench (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout -s KILL 900 python -m pytest tests/test_gpu_decomp.py -x -q > gpurun_out/pytest_decomp.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_decomp.log
for mode in nccl p2p; do
  LJ_HALO=$mode timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "bench $mode rc=$?"; cat gpurun_out/bench_n${N}_$mode.json; tail -3 gpurun_out/bench_n${N}_$mode.err
done
