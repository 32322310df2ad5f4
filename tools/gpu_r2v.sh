set -x
ncu --set full --clock-control none --import-source on -k regex:scan_mma256w -s 4 -c 1 -o gpurun_out/r2v_c2_main -f python bench.py --workload c2 --no-cpu-baseline --sweep '' --threads 0 --steps 2 --warmup 3 --no-parity > gpurun_out/r2v_ncu_c2.log 2>&1
ncu -i gpurun_out/r2v_c2_main.ncu-rep --page raw --csv > gpurun_out/r2v_c2_main_raw.csv 2>/dev/null
ncu -i gpurun_out/r2v_c2_main.ncu-rep --page details > gpurun_out/r2v_c2_main_details.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 4 -c 1 -o gpurun_out/r2v_scan_mma_b64 -f python bench.py --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 2 --warmup 3 --no-parity > gpurun_out/r2v_ncu_b64.log 2>&1
ncu -i gpurun_out/r2v_scan_mma_b64.ncu-rep --page raw --csv > gpurun_out/r2v_scan_mma_b64_raw.csv 2>/dev/null
ncu -i gpurun_out/r2v_scan_mma_b64.ncu-rep --page details > gpurun_out/r2v_scan_mma_b64_details.txt 2>/dev/null
ls -la gpurun_out/r2v_*
