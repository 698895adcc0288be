#!/bin/bash
# round 2, step I (2 GPUs): full GPU suite incl. the 2-GPU parity tests, rows leg, weak / strong lines at N = 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_2gpu.log
tail -8 gpurun_out/pytest_gpu_2gpu.log | cut -c1-300
timeout 600 python bench_tools/rows_leg.py
bash bench_tools/gpu_r2m.sh r2i "2" "n2_1e6:weak n2_1e6:strong li2o_1e5:strong li2o_1e5:weak"
