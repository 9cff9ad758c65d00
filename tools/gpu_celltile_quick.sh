#!/bin/bash
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "celltile or six_array or decomp" 2>&1 | tail -2
timeout -s KILL 300 python tools/celltile_check.py --reps 20 2>&1 | grep -E "mixed|force:|rror"
