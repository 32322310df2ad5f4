set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "rc=$?" >> gpurun_out/r2a_bench_n1.err
python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 100 > gpurun_out/r2a_bench_shard1250k.json 2> gpurun_out/r2a_bench_shard1250k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches_shard1250k.csv python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 10 --no-parity > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; tail -c 600 gpurun_out/r2a_bench_n1.json; tail -c 400 gpurun_out/r2a_bench_shard1250k.json
