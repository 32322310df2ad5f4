set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ksplit or default_dispatch" > gpurun_out/r2e_pytest_ks.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_ks.log
tail -3 gpurun_out/r2e_pytest_ks.log
timeout 600 python bench.py --rows 6250000 --dim 1536 --batch 64 --sweep '1,16' --threads 0 --steps 20 --no-parity > gpurun_out/r2e_bench_1536_b64.json 2> gpurun_out/r2e_bench_1536_b64.err; echo "rc=$?"
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_golden.py tests/test_hybrid.py -x -q -m gpu > gpurun_out/r2e_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_multi.log
tail -15 gpurun_out/r2e_pytest_multi.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],3), [(x['batch'],round(x['qps']),round(x['scan_frac_of_hbm_peak'],3)) for x in d['sweep']])
    except Exception as e: print(f,'ERR',e)
PY
