#!/bin/bash
# ncu --set full capture of the fused kernel for a workload:  bash bench_tools/gpu_ncu.sh <workload> <tag> [kernel regex]
WL=${1:-n2_1e6}; TAG=${2:-prof}; KR=${3:-eloc_sliced}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -s 3 -c 1 -o gpurun_out/$TAG \
   python bench.py --steps 2 --warmup 3 --workload $WL --cpu-sample 0 --no-e2e > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/$TAG.ncu-rep
