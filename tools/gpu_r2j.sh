set -x
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest_gpu_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest_gpu_full.log
tail -8 gpurun_out/r2j_pytest_gpu_full.log
timeout 600 python bench.py --workload c5 > gpurun_out/r2j_bench_c5.json 2> gpurun_out/r2j_bench_c5.err; echo "rc=$?"
timeout 600 python bench.py --workload c2 > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench_c2.err; echo "rc=$?"
timeout 600 python bench.py --workload c3 --threads 0 > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err; echo "rc=$?"
timeout 600 python bench.py --workload c3 --batch 1024 --sweep '' --threads 0 --steps 10 > gpurun_out/r2j_bench_c3_b1024.json 2> gpurun_out/r2j_bench_c3_b1024.err; echo "rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), round(d['ms_per_step'],4), (d.get('roofline') or {}).get('kernel'), round((d.get('roofline') or {}).get('frac',0),3), (d.get('parity') or {}).get('status'), [(x['batch'], round(x['qps']), x['kernel']) for x in d.get('sweep',[])])
    except Exception as e: print(f,'ERR',e)
PY
