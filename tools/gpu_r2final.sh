set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest_gpu_full.log
tail -3 gpurun_out/r2f_pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench_n1_default.json 2> gpurun_out/r2f_bench_n1_default.err; echo "rc=$?"
timeout 600 python bench.py --workload c2 --sweep '' --threads 0 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "rc=$?"
timeout 600 python bench.py --workload c3 --sweep '' --threads 0 > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err; echo "rc=$?"
timeout 600 python bench.py --workload c5 > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err; echo "rc=$?"
timeout 600 python bench.py --rows 6250000 --dim 1536 --no-cpu-baseline --sweep '1,16' --threads 0 --also-f32 0 --steps 30 > gpurun_out/r2f_bench_6250k_1536_b64.json 2>/dev/null; echo "rc=$?"
timeout 600 python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 > gpurun_out/r2f_bench_shard1250k_b64.json 2>/dev/null; echo "rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches_bench_b64_k10.csv python bench.py --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 3 --warmup 3 --no-parity > gpurun_out/r2f_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:(round(v,4) if v else v) for k,v in d['phases_ms'].items()}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['roofline']['frac'],3), round(d['e2e']['value'],1))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/r2f_smoke.log | tail -4
