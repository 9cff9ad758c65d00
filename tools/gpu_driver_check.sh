#!/bin/bash
# the C++ driver on a GPU box: goldens for the CUDA program and the OpenACC SoA program, and the
# bench line in mixed precision
timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "cpp_driver" 2>&1 | tail -3
timeout -s KILL 300 python bench.py --prec mixed --steps 100 --warmup 20 --no-cpu 2>&1 | tail -1 | cut -c1-1500
