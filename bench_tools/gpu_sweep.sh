#!/bin/bash
# BASELINE config 5: synthetic Pauli-sum sweep, 64- and 128-bit masks (device-resident couplings/s; 20 steps each)
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for cfg in "64 1000 100000" "64 10000 100000" "64 10000 1000000" "64 100000 100000" "64 100000 1000000" "127 10000 100000" "127 10000 1000000" "127 100000 100000" "40 10000 1000000"; do
  set -- $cfg
  timeout 600 python bench.py --workload synthetic --synthetic $1 $2 $3 --steps 20 --warmup 3 --cpu-sample 0 --no-extras --no-e2e >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
  tail -n 1 gpurun_out/sweep.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N K M = $1 $2 $3 : value %.3e  ms/step %.3f  kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
done
