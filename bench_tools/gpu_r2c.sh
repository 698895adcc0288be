#!/bin/bash
# round 2, step C: parity suite + Li2O bench with the record-mask hash walk (A/B: bank binning off) + ncu of the Li2O kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
run() {  # name, env..., workload
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --cpu-sample 0 --no-extras > gpurun_out/r2c_${name}.json 2> gpurun_out/r2c_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_${name}.json").read().strip().splitlines()[-1])
    print("${name} value %.3e kernel_ms %.4f ms_per_step %.4f e2e %.3e launches %d check %s" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["check"]))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2c_${name}.err").read()[-2000:])
PY
}
run li2o_bin NAQS_X=1 BENCH_WL=li2o_1e5
run li2o_nobin NAQS_ELOC_NO_BIN=1 BENCH_WL=li2o_1e5
run n2 NAQS_X=1
bash bench_tools/gpu_ncu.sh li2o_1e5 r02c_li2o "eloc_sliced"
