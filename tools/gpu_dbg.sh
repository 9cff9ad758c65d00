#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "six_array or decomp" 2>&1 | tail -3
for n in 1 2; do
    if [ $n -eq 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py"; fi
    timeout -s KILL 600 $CMD --gpus $n --steps 100 --warmup 20 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "bench n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
    print("n=%d value=%.4g ms/step=%.4f launches=%s halo=%s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["gpu_launches"],d.get("halo")))
except Exception as e:
    print("parse failed",e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
