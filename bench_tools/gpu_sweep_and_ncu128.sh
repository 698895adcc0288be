#!/bin/bash
# (1 GPU) BASELINE config 5 grid in one process + ncu --set full of the 128-bit hash walk
mkdir -p gpurun_out
timeout 1500 python bench_tools/sweep.py --out gpurun_out/sweep.jsonl 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eloc_sliced -s 3 -c 1 -o gpurun_out/synth127 \
   python bench.py --steps 2 --warmup 3 --workload synthetic --synthetic 127 10000 100000 --cpu-sample 0 --no-e2e --no-extras > gpurun_out/synth127_ncu.log 2>&1
ls -la gpurun_out/synth127.ncu-rep
