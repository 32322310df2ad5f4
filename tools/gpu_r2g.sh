set -x
timeout 1500 python -m pytest tests/test_gpu_pool.py tests/test_hybrid.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2g_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_new.log
tail -15 gpurun_out/r2g_pytest_new.log
# ncu: full capture of the k-split kernel's main launch (6.25M x 1536 bf16, 64 queries), after a plain run exited 0
timeout 300 python tools/one_search.py 64 10 3 6250000 bf16 1536 > gpurun_out/r2g_one_ks.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -s 3 -c 1 -o gpurun_out/r2g_scan_mma_ks_b64 -f python tools/one_search.py 64 10 2 6250000 bf16 1536 > gpurun_out/r2g_ncu_ks.log 2>&1
ncu -i gpurun_out/r2g_scan_mma_ks_b64.ncu-rep --page raw --csv > gpurun_out/r2g_scan_mma_ks_bf16_1536_b64_ncu_full.csv 2>/dev/null
# launch list of one dim-1536 step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_1536_b64.csv python tools/one_search.py 64 10 3 6250000 bf16 1536 > gpurun_out/r2g_ncu_ll.log 2>&1
tail -3 gpurun_out/r2g_one_ks.log
