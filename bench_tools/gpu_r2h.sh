#!/bin/bash
# round 2, step H: full GPU suite (CUDA-graph small batches, fused-H bound) + LiH call latency + launch list of the rows leg
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench_tools/rows_leg.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rows_launches.csv python bench_tools/rows_leg.py > gpurun_out/rows_ncu.log 2>&1
python bench_tools/launch_list_md.py gpurun_out/rows_launches.csv "rows leg" | head -30
timeout 900 python bench.py --steps 30 --warmup 5 --no-vmc --cpu-sample 0 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
oc=d["other_configs"]
for k,v in oc.items(): print(k,{a:b for a,b in v.items() if a!="workload"})
PY
NAQS_ELOC_NO_GRAPH=1 timeout 900 python bench.py --steps 10 --warmup 5 --no-vmc --cpu-sample 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-graph LiH call', d['other_configs']['lih_vmc_eloc_call'])"
