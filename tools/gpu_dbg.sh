#!/bin/bash
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/debug/slab_time.py 2>&1 | grep "rank\|rror" | head
LJ_HALO=nccl timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/debug/slab_time.py 2>&1 | grep "rank\|rror" | head
