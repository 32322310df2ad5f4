set -x
ncu --set full --clock-control none --import-source on -k regex:"hybrid_mask_kernel|hybrid_pair_score" -s 6 -c 2 -o gpurun_out/r2z2_c5 -f python bench.py --workload c5 --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2z2_ncu_c5.log 2>&1
ncu -i gpurun_out/r2z2_c5.ncu-rep --page raw --csv > gpurun_out/r2z2_c5_raw.csv 2>/dev/null
ncu -i gpurun_out/r2z2_c5.ncu-rep --page details > gpurun_out/r2z2_c5_details.txt 2>/dev/null
ncu -i gpurun_out/r2z2_c5.ncu-rep --page source --csv > gpurun_out/r2z2_c5_source.csv 2>/dev/null
