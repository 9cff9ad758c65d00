#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile or decomp or row_range" 2>&1 | tail -3
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/debug/slab_time.py 2>&1 | grep "rank\|rror" | head -4
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 20 --no-cpu > gpurun_out/scale_n2b.json 2> gpurun_out/scale_n2b.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/scale_n2b.json
