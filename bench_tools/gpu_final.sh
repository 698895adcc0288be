#!/bin/bash
# (1 GPU) final suite + bench lines + launch list (the ncu --set full captures of the unchanged hot kernels are those of step F)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 600 gpurun_out/final_bench.json
for wl in li2o_1e5 h2o_1e5; do
  BENCH_WL=$wl timeout 600 python bench.py --steps 30 --warmup 5 --cpu-sample 0 --no-extras > gpurun_out/r2o_$wl.json 2> gpurun_out/r2o_$wl.err
  tail -c 300 gpurun_out/r2o_$wl.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --steps 10 --warmup 5 --cpu-sample 0 --no-e2e --no-extras > gpurun_out/final_launches.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/final_smi.csv
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
tail -c 400 gpurun_out/final_bench_reference.json
python __graft_entry__.py smoke 2>&1 | tail -2
