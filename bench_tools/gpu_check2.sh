#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
./bench_tools/pipe_peaks > gpurun_out/pipe_peaks.json 2>&1
cat gpurun_out/pipe_peaks.json
for wl in n2_1e6 li2o_1e5 h2o_1e5; do
  timeout 600 python bench.py --steps 30 --warmup 5 --workload $wl --cpu-sample 0 > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}.json").read().strip().splitlines()[-1])
print("$wl sliced value %.3e kernel_ms %.4f ms_per_step %.4f e2e %.3e launches %d" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
  NAQS_ELOC_ALGO=direct timeout 600 python bench.py --steps 30 --warmup 5 --workload $wl --cpu-sample 0 --no-e2e > gpurun_out/bench_${wl}_direct.json 2> gpurun_out/bench_${wl}_direct.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_direct.json").read().strip().splitlines()[-1])
print("$wl direct value %.3e kernel_ms %.4f ms_per_step %.4f" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"]))
PY
done
tail -n 5 gpurun_out/*.err
