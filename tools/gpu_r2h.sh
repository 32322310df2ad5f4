set -x
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_sharded.py "tests/test_gpu_multi.py::test_spread_over_all_visible_gpus" -x -q -m gpu > gpurun_out/r2h_pytest_multi_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest_multi_gpu.log
tail -5 gpurun_out/r2h_pytest_multi_gpu.log
timeout 600 $TR bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err; echo "rc=$?"
timeout 900 $TR bench.py --gpus 8 --workload c4 --steps 50 > gpurun_out/r2h_bench_c4_n8_b1.json 2> gpurun_out/r2h_bench_c4_n8_b1.err; echo "rc=$?"
timeout 900 $TR bench.py --gpus 8 --workload c4 --batch 64 --steps 50 --no-parity > gpurun_out/r2h_bench_c4_n8_b64.json 2> gpurun_out/r2h_bench_c4_n8_b64.err; echo "rc=$?"
timeout 900 python tools/multi_c4.py 6250000 20 > gpurun_out/r2h_multi_c4.json 2> gpurun_out/r2h_multi_c4.err; echo "rc=$?"
tail -c 700 gpurun_out/r2h_bench_n8.json; tail -c 500 gpurun_out/r2h_multi_c4.json; tail -5 gpurun_out/r2h_multi_c4.err
