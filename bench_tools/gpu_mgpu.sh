#!/bin/bash
# multi-GPU step: 2-GPU parity test (C-ABI exchange included) + weak / strong bench lines at every N <= visible GPUs.
# usage: bash bench_tools/gpu_mgpu.sh <tag> [workloads...]   (default workloads: n2_1e6 li2o_1e5)
TAG=${1:-mgpu}; shift
WLS=${@:-"n2_1e6 li2o_1e5"}
NG=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_gpus" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -25 gpurun_out/pytest_${TAG}.log | cut -c1-250
for wl in $WLS; do
 for mode in weak strong; do
  for n in 1 2 4 8; do
   [ $n -gt $NG ] && break
   [ "$mode" = "strong" ] && [ $n = 1 ] && continue
   if [ "$n" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py"; fi
   FLAG=""; [ "$mode" = "strong" ] && FLAG="--strong"
   out=gpurun_out/${TAG}_${wl}_${mode}_g$n
   timeout 900 $CMD --gpus $n --steps 50 --warmup 5 --cpu-sample 0 --no-extras --workload $wl $FLAG > $out.json 2> $out.err
   python - <<PY
import json
try:
    d=json.loads(open("$out.json").read().strip().splitlines()[-1])
    print("$wl $mode gpus $n value %.3e ms_per_step %.4f kernel_ms %.4f e2e %.3e check %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"] if d["e2e"] else 0, d["check"].get("multi_gpu_vs_single_rank")))
except Exception as e:
    print("$wl $mode gpus $n FAILED", e); print(open("$out.err").read()[-2500:])
PY
  done
 done
done
