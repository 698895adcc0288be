#!/bin/bash
# final evidence for profiles/: launch list of the default bench command + ncu --set full of the hot kernel (N2, Li2O)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --steps 10 --warmup 5 --cpu-sample 0 --no-e2e --no-extras > gpurun_out/final_launches.log 2>&1
bash bench_tools/gpu_ncu.sh n2_1e6 final_n2
bash bench_tools/gpu_ncu.sh li2o_1e5 final_li2o
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/final_smi.csv
python bench.py --steps 100 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
tail -c 400 gpurun_out/final_bench.json
