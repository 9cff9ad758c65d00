#!/bin/bash
# weak scaling refresh at 4 and 8 GPUs (1 and 2 are measured by tools/gpu_multi.sh)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
for n in 4 8; do
  if [ $n -le $N ]; then
    timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 20 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "bench n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
    print("n=%d value=%.4g ms/step=%.4f launches=%s halo=%s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["gpu_launches"],d.get("halo")))
except Exception as e:
    print("parse failed",e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
  fi
done
