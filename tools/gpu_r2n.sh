set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log
tail -3 gpurun_out/r2n_pytest.log
for rep in 1 2; do
for g in 0 2; do
for b in 128 256 1024; do
MRAG_GMAX=$g timeout 300 python bench.py --batch $b --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 20 --no-parity > gpurun_out/r2n_b${b}_g${g}_$rep.json 2>/dev/null
done; done; done
for g in 0 2; do
MRAG_GMAX=$g timeout 300 python bench.py --workload c2 --no-cpu-baseline --sweep '' --threads 0 > gpurun_out/r2n_c2_g${g}.json 2>/dev/null
MRAG_GMAX=$g timeout 300 python bench.py --dtype f32 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 20 > gpurun_out/r2n_f32_g${g}.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
