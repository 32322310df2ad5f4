set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "mma128 or pair or approx or rescore or fallback or c2 or c3 or full or large or candidate" > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_pytest.log
tail -3 gpurun_out/r2w_pytest.log
for rep in 1 2; do
timeout 300 python bench.py --workload c2 --no-cpu-baseline --sweep '' --threads 0 --no-parity > gpurun_out/r2w_c2_$rep.json 2>/dev/null
for b in 256 1024; do
timeout 300 python bench.py --batch $b --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 20 --no-parity > gpurun_out/r2w_b${b}_$rep.json 2>/dev/null
done
done
timeout 300 python bench.py --workload c3 --batch 1024 --sweep '' --threads 0 --steps 10 --no-parity --no-cpu-baseline > gpurun_out/r2w_c3_b1024.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2w_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items() if v}, d['gpu_launches'], round(d['roofline']['frac'],3), d['roofline'].get('tensor',{}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
