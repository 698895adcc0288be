#!/bin/bash
# round 2, step J (2 GPUs): 2-GPU parity tests with the merge exchange + N2 weak/strong with merge vs NCCL all-reduce
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "two_gpus" > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2j.log
tail -8 gpurun_out/pytest_r2j.log | cut -c1-300
bash bench_tools/gpu_r2m.sh r2j_merge "2" "n2_1e6:weak n2_1e6:strong"
NAQS_BENCH_EXCHANGE_FLAGS=0x4000 bash bench_tools/gpu_r2m.sh r2j_nccl "2" "n2_1e6:weak"
