#!/bin/bash
# round 2, step A: parity suite + N2 / H2O bench with the key-order kernel v3 and (A/B) the generic key-order walk
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
run() {  # name, env..., workload
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --cpu-sample 0 --no-extras > gpurun_out/r2a_${name}.json 2> gpurun_out/r2a_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_${name}.json").read().strip().splitlines()[-1])
    print("${name} value %.3e kernel_ms %.4f ms_per_step %.4f e2e %.3e launches %d check %s" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["check"]))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2a_${name}.err").read()[-2000:])
PY
}
run n2_ko3 NAQS_X=1
run n2_old NAQS_ELOC_NO_KO3=1
run h2o_ko3 NAQS_X=1 BENCH_WL=h2o_1e5
run h2o_old NAQS_ELOC_NO_KO3=1 BENCH_WL=h2o_1e5
run li2o NAQS_X=1 BENCH_WL=li2o_1e5
bash bench_tools/gpu_ncu.sh n2_1e6 r02a_n2 "eloc_keyorder|eloc_sliced"
