#!/bin/bash
timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "cpp_driver" 2>&1 | tail -3
lj_gpu_b200/driver/force_b200 --soa6 2>&1 >/dev/null | head -3
