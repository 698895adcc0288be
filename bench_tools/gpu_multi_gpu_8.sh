#!/bin/bash
# (8 GPUs) exchange probe + weak / strong lines at N = 8 (merge vs NCCL all-reduce for the dense-shard exchange)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29801 bench_tools/exchange_probe.py 2>/dev/null | tail -1 | tee gpurun_out/r2k_exchange_probe_g$NG.json
run() {  # tag n workload mode [env]
  local out=gpurun_out/r2k_$1_$3_$4_g$2
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
run merge $NG n2_1e6 weak
NAQS_BENCH_EXCHANGE_FLAGS=0x4000 run nccl $NG n2_1e6 weak
run merge $NG n2_1e6 strong
run auto $NG li2o_1e5 strong
run auto $NG li2o_1e5 weak
run merge 4 n2_1e6 weak
