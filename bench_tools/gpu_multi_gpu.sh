#!/bin/bash
# round 2, multi-GPU step: [2-GPU parity tests] + weak / strong bench lines for the N given (bench self-checks against a single-rank recomputation)
# usage: bash bench_tools/gpu_multi_gpu.sh <tag> "<N list>" "<workload:mode list>" [pytest]
TAG=${1:-r2m}; NS=${2:-"2"}; JOBS=${3:-"n2_1e6:weak n2_1e6:strong li2o_1e5:strong"}; PYT=${4:-}
NG=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
if [ -n "$PYT" ]; then
  timeout 600 python -m pytest tests -q -m gpu -k "two_gpus" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
  tail -8 gpurun_out/pytest_${TAG}.log | cut -c1-250
fi
port=29610
for job in $JOBS; do
 wl=${job%%:*}; mode=${job##*:}
 for n in $NS; do
   [ $n -gt $NG ] && continue
   port=$((port+1))
   FLAG=""; [ "$mode" = "strong" ] && FLAG="--strong"
   out=gpurun_out/${TAG}_${wl}_${mode}_g$n
   timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 50 --warmup 5 --cpu-sample 0 --no-extras --workload $wl $FLAG > $out.json 2> $out.err
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
