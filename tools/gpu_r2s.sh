set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log
tail -3 gpurun_out/r2s_pytest.log
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2 3; do
eval timeout 300 $B > gpurun_out/r2s_shard_$rep.json 2>/dev/null
done
timeout 300 python bench.py --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 30 > gpurun_out/r2s_10m.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
