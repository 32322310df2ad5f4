set -x
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pool.py tests/test_hybrid.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2f_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest_new.log
tail -25 gpurun_out/r2f_pytest_new.log
timeout 600 python bench.py --rows 6250000 --dim 1536 --batch 64 --sweep '' --threads 0 --steps 40 --no-parity --no-cpu-baseline > gpurun_out/r2f_1536_b64.json 2> gpurun_out/r2f_1536_b64.err
timeout 600 python bench.py --rows 6250000 --dim 1536 --batch 1 --sweep '' --threads 0 --steps 40 --no-parity --no-cpu-baseline > gpurun_out/r2f_1536_b1.json 2> gpurun_out/r2f_1536_b1.err
timeout 600 python bench.py --rows 6250000 --dim 1536 --batch 16 --sweep '' --threads 0 --steps 40 --no-parity --no-cpu-baseline > gpurun_out/r2f_1536_b16.json 2> gpurun_out/r2f_1536_b16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
