#!/bin/bash
# (1 GPU) final suite after the warp-reconvergence points + N2 / Li2O / synthetic-127 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
for wl in n2_1e6 li2o_1e5; do
  timeout 300 python bench.py --steps 50 --warmup 5 --cpu-sample 0 --no-extras --no-e2e --workload $wl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl value %.3e ms/step %.4f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
done
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --no-extras --no-e2e --workload synthetic --synthetic 127 10000 100000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('synthetic 127 1e4 1e5 value %.3e ms/step %.4f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
