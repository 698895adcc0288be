#!/bin/bash
# round 2, step G: full GPU suite (device-resident loop fix, fused-H test, chunked rows) + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 100 --warmup 5 --no-vmc > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.3e ms/step %.4f kernel_ms %.4f e2e %.3e serial %.3e frac %.3f bound %s"%(d["value"],d["ms_per_step"],r["kernel_ms"],d["e2e"]["value"],d["e2e"]["pipeline"]["serial_value"],r["frac"],r["bound"]))
oc=d["other_configs"]
for k,v in oc.items(): print(k,{a:b for a,b in v.items() if a!="workload"})
PY
tail -5 gpurun_out/r2g_bench.err
