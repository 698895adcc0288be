#!/bin/bash
# (8 GPUs) push-gather exchange at N = 8 (Li2O strong / weak), N2 weak at 2 and strong at 4 with the final merge kernel
mkdir -p gpurun_out
run() {  # tag n workload mode
  local out=gpurun_out/r2p_$1_$3_$4_g$2
  FLAG=""; [ "$4" = "strong" ] && FLAG="--strong"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29810 + RANDOM % 100)) bench.py --gpus $2 --steps 40 --warmup 5 --cpu-sample 0 --no-extras --no-e2e --workload $3 $FLAG > $out.json 2> $out.err
  python - <<PY
import json
try:
    d=json.loads(open("$out.json").read().strip().splitlines()[-1])
    print("$1 $3 $4 gpus $2 value %.3e ms_per_step %.4f kernel_ms %.4f check_ok %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], (d["check"].get("multi_gpu_vs_single_rank") or {}).get("ok")))
except Exception as e:
    print("$1 $3 $4 gpus $2 FAILED", e); print(open("$out.err").read()[-1500:])
PY
}
run pushgather 8 li2o_1e5 strong
run pushgather 8 li2o_1e5 weak
run merge 2 n2_1e6 weak &
CUDA_VISIBLE_DEVICES=2,3,4,5 run merge 4 n2_1e6 strong &
wait
