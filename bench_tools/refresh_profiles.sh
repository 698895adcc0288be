#!/bin/bash
# copy the outputs of bench_tools/gpu_profile_final.sh from gpurun_out/ into profiles/ (round tag $1, default r01) and rebuild the summaries
R=${1:-r01}
ncu -i gpurun_out/final_n2.ncu-rep --page raw --csv > profiles/${R}_final_n2_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/final_li2o.ncu-rep --page raw --csv > profiles/${R}_final_li2o_ncu_raw.csv 2>/dev/null
python bench_tools/ncu_summarize.py profiles/ncu_summary_${R}.json n2_1e6=profiles/${R}_final_n2_ncu_raw.csv li2o_1e5=profiles/${R}_final_li2o_ncu_raw.csv > /dev/null
cp gpurun_out/final_launches.csv profiles/${R}_final_launches_n2_1e6.csv
cp gpurun_out/final_bench.json profiles/${R}_final_bench_n2_1e6.json
cp gpurun_out/final_bench_reference.json profiles/${R}_final_bench_reference.json
cp gpurun_out/final_smi.csv profiles/${R}_final_smi.csv
python bench_tools/launch_list_md.py profiles/${R}_final_launches_n2_1e6.csv "round ${R#r} final, \`python bench.py --steps 10 --warmup 5 --cpu-sample 0 --no-e2e --no-extras\` (N2, M = 1e6, 1 B200)" > profiles/${R}_final_launches_n2_1e6.md
python bench_tools/make_profiles_readme.py ${R} > /dev/null
