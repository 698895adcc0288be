#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
  compute-sanitizer --tool memcheck python -c "import torch; print(torch.zeros(4, device='cuda').sum().item())" > gpurun_out/memprobe_warm.log 2>&1 && break
  tail -2 gpurun_out/memprobe_warm.log | cut -c1-160; sleep 5
done
tail -1 gpurun_out/memprobe_warm.log | cut -c1-100
for v in plain nobin static; do
  unset NAQS_ELOC_NO_BIN NAQS_ELOC_STATIC_TASKS
  [ $v = nobin ] && export NAQS_ELOC_NO_BIN=1
  [ $v = static ] && export NAQS_ELOC_STATIC_TASKS=1
  for try in 1 2 3; do
    timeout 400 compute-sanitizer --tool memcheck --print-limit 4 --show-backtrace device python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_batches_with_global_filter and 70-400" > gpurun_out/memprobe_$v.log 2>&1
    grep -q "before first instrumented" gpurun_out/memprobe_$v.log || break
  done
  echo "== $v (try $try): $(grep -c 'Invalid' gpurun_out/memprobe_$v.log) invalid; $(grep 'ERROR SUMMARY' gpurun_out/memprobe_$v.log)"; grep -m8 "Invalid\|Access at\|by thread\|cuh:" gpurun_out/memprobe_$v.log | cut -c1-200; tail -3 gpurun_out/memprobe_$v.log | cut -c1-200
done
