#!/bin/bash
# round 2, step L (2 GPUs): push-gather exchange of the hash path — 2-GPU parity tests, Li2O fixture, Li2O strong / weak at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "two_gpus or Li2O_500" > gpurun_out/pytest_r2l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2l.log
tail -8 gpurun_out/pytest_r2l.log | cut -c1-300
bash bench_tools/gpu_r2m.sh r2l "2" "li2o_1e5:strong li2o_1e5:weak"
