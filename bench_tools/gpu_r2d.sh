#!/bin/bash
# round 2, step D: Li2O chunk / dealing / binning sweep + VMC tests (fused loss, device-resident)
mkdir -p gpurun_out
run() {  # name, env..., workload
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --no-extras --no-e2e > gpurun_out/r2d_${name}.json 2> gpurun_out/r2d_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_${name}.json").read().strip().splitlines()[-1])
    print("${name} value %.3e kernel_ms %.4f ms_per_step %.4f check %s" % (d["value"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["check"]))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2d_${name}.err").read()[-1500:])
PY
}
for ch in 3 4 5 7 10 14; do
  run li2o_bin_dyn_c$ch NAQS_ELOC_CHUNKS=$ch BENCH_WL=li2o_1e5
  run li2o_nobin_dyn_c$ch NAQS_ELOC_CHUNKS=$ch NAQS_ELOC_NO_BIN=1 BENCH_WL=li2o_1e5
done
run li2o_bin_static_c7 NAQS_ELOC_CHUNKS=7 NAQS_ELOC_STATIC_TASKS=1 BENCH_WL=li2o_1e5
run li2o_default NAQS_X=1 BENCH_WL=li2o_1e5
run n2_default NAQS_X=1
run h2o_default NAQS_X=1 BENCH_WL=h2o_1e5
timeout 1200 python -m pytest tests/test_vmc_gpu.py -q -m gpu > gpurun_out/pytest_vmc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_vmc.log
tail -40 gpurun_out/pytest_vmc.log | cut -c1-250
