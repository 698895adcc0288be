#!/bin/bash
# round 2, step E: device API / fused loss / range-check tests + VMC tests + full suite + bench line with the live roofline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-250
for wl in n2_1e6 li2o_1e5 h2o_1e5; do
  BENCH_WL=$wl timeout 600 python bench.py --steps 30 --warmup 5 --cpu-sample 0 --no-extras > gpurun_out/r2e_$wl.json 2> gpurun_out/r2e_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_$wl.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$wl value %.3e kernel_ms %.4f e2e %.3e roofline %s frac %.3f t_issue %.3f t_l1 %.3f" % (d["value"], r["kernel_ms"], d["e2e"]["value"], r["bound"], r["frac"], r["model"]["t_issue_ms"], r["model"]["t_l1tex_ms"]))
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/r2e_$wl.err").read()[-1500:])
PY
done
