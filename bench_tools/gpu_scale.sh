#!/bin/bash
# weak-scaling run exactly as the driver launches it: N = 1, 2, 4, 8 (as many as visible)
NG=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ "$n" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py"; fi
  timeout 900 $CMD --gpus $n --steps 50 --warmup 5 --cpu-sample 0 > gpurun_out/scale_g$n.json 2> gpurun_out/scale_g$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_g$n.json").read().strip().splitlines()[-1])
    print("gpus $n value %.3e ms_per_step %.4f kernel_ms %.4f e2e %.3e clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("gpus $n FAILED", e); print(open("gpurun_out/scale_g$n.err").read()[-3000:])
PY
done
