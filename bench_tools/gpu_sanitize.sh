#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a parity subset that reaches every kernel family of round 2:
# sliced / key-order / hash walks, chunked rows, direct kernel with chunks, CUDA-graph replay, loss terms, Level-0 twins
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout 1500 compute-sanitizer --tool $tool --launch-timeout 0 --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_device_api_gpu.py tests/test_fused_h_gpu.py -x -q -m gpu \
     -k "fixture and (LiH_small or LiH_full or N2_2000 or H2O_300_f32 or Li2O_500) or edge_cases or wide_and_large_synthetic_tables and 40 or full_size_li2o or boundary_widths or large_batches or caller_owned or (rows_and_hij and LiH) or level0 or replay_a_cuda_graph or loss_terms or out_of_range or (fused_matrix and (LiH or 16384-40))" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
  tail -6 gpurun_out/sanitize_$tool.log
done
