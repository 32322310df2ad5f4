set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_gpu_full.log
tail -3 gpurun_out/r2g_pytest_gpu_full.log
timeout 600 python bench.py --workload c5 > gpurun_out/r2g_bench_c5.json 2> gpurun_out/r2g_bench_c5.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_c5.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], d['roofline']['kernel'], d['cpu_baseline'])
PY
