#!/bin/bash
# One gpurun call: smoke + parity tests + pipe ceilings + bench + ncu launch list.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
./bench_tools/pipe_peaks > gpurun_out/pipe_peaks.json 2>&1
cat gpurun_out/pipe_peaks.json
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 3000 gpurun_out/bench_n2.json
timeout 600 python bench.py --steps 50 --warmup 5 --workload li2o_1e5 --cpu-sample 20000 > gpurun_out/bench_li2o.json 2> gpurun_out/bench_li2o.err; tail -c 1500 gpurun_out/bench_li2o.json
timeout 300 python bench.py --steps 50 --warmup 5 --workload h2o_1e5 --cpu-sample 0 > gpurun_out/bench_h2o.json 2> gpurun_out/bench_h2o.err; tail -c 1500 gpurun_out/bench_h2o.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/*.err
