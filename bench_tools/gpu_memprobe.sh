#!/bin/bash
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --launch-timeout 0 --print-limit 3 python bench_tools/memcheck_probe.py 2 > gpurun_out/memprobe_dyn.log 2>&1
grep -n "Invalid\|by thread\|Access at\|rep \|ERROR SUMMARY" gpurun_out/memprobe_dyn.log | cut -c1-150 | head -12
TOOLS=memcheck bash bench_tools/gpu_sanitize.sh
