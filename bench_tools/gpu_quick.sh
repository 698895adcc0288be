#!/bin/bash
# quick iteration: parity tests + the three workloads (sliced only)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in ${WORKLOADS:-n2_1e6 li2o_1e5 h2o_1e5}; do
  timeout 600 python bench.py --steps 30 --warmup 5 --workload $wl --cpu-sample 0 > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${wl}.json").read().strip().splitlines()[-1])
    print("$wl value %.3e kernel_ms %.4f ms_per_step %.4f e2e %.3e launches %d" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/bench_${wl}.err").read()[-2000:])
PY
done
