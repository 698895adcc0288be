#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpus or edge_cases" 2>&1 | tail -3
for n in 1 $N; do
  if [ "$n" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py"; fi
  timeout 900 $CMD --gpus $n --steps 50 --warmup 5 --cpu-sample 0 --no-extras > gpurun_out/scale2_g$n.json 2> gpurun_out/scale2_g$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale2_g$n.json").read().strip().splitlines()[-1])
    print("gpus $n value %.3e ms_per_step %.4f kernel_ms %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"]))
except Exception as e:
    print("gpus $n FAILED", e); print(open("gpurun_out/scale2_g$n.err").read()[-3000:])
PY
done
