set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ksplit or default_dispatch" > gpurun_out/r2b_pytest_ks.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest_ks.log
tail -5 gpurun_out/r2b_pytest_ks.log
timeout 600 python bench.py --rows 6250000 --dim 1536 --batch 64 --sweep '1,16' --threads 0 --steps 20 > gpurun_out/r2b_bench_1536_b64.json 2> gpurun_out/r2b_bench_1536_b64.err; echo "rc=$?"
timeout 600 python bench.py --workload c4 --threads 0 > gpurun_out/r2b_bench_c4_b1.json 2> gpurun_out/r2b_bench_c4_b1.err; echo "rc=$?"
timeout 600 python bench.py --workload c4 --batch 64 --sweep '' --threads 0 > gpurun_out/r2b_bench_c4_b64.json 2> gpurun_out/r2b_bench_c4_b64.err; echo "rc=$?"
MRAG_LIB=$PWD/build_ab/libmrag_stamps.so timeout 300 python tools/stats_probe.py 1250000 10 > gpurun_out/r2b_stamps_1250k.txt 2>&1
tail -c 1500 gpurun_out/r2b_stamps_1250k.txt
