#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$1" == "D" ]; then
  timeout -s KILL 1200 python tools/run_configs.py --config D > gpurun_out/config_D.jsonl 2> gpurun_out/config_D.err; echo "D rc=$?"; cat gpurun_out/config_D.jsonl; tail -3 gpurun_out/config_D.err
elif [ "$1" == "E1" ]; then
  timeout -s KILL 1500 python tools/run_configs.py --config E1 > gpurun_out/config_E1.jsonl 2> gpurun_out/config_E1.err; echo "E1 rc=$?"; cat gpurun_out/config_E1.jsonl; tail -3 gpurun_out/config_E1.err
elif [ "$1" == "E" ]; then
  LJ_BENCH_CELLS=320 timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --density 0.8 --steps 40 --warmup 20 --no-cpu > gpurun_out/config_E$N.json 2> gpurun_out/config_E$N.err; echo "E$N rc=$?"; cat gpurun_out/config_E$N.json; tail -3 gpurun_out/config_E$N.err
fi
