#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small parity subset
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
     -k "fixture and (LiH_small or LiH_full or N2_2000 or H2O_300_f32) or edge_cases or wide_and_large_synthetic_tables and 40 or full_size_li2o or boundary_widths or large_batches or caller_owned" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
  tail -6 gpurun_out/sanitize_$tool.log
done
