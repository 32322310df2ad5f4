set -x
for rep in 1 2; do
for mk in 16 128; do
MRAG_GMAX_MAXK=$mk timeout 300 python bench.py --workload c3 --sweep '' --threads 0 --steps 20 --no-cpu-baseline > gpurun_out/r2x_c3_b64_mk${mk}_$rep.json 2>/dev/null
MRAG_GMAX_MAXK=$mk timeout 300 python bench.py --workload c3 --rows 1250000 --sweep '' --threads 0 --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2x_c3_shard_mk${mk}_$rep.json 2>/dev/null
done
for b in 256 1024; do
timeout 300 python bench.py --batch $b --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 20 --no-parity > gpurun_out/r2x_b${b}_$rep.json 2>/dev/null
done
done
timeout 300 python bench.py --workload c2 --no-cpu-baseline --sweep '' --threads 0 > gpurun_out/r2x_c2.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items() if v}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['roofline']['frac'],3), d['roofline'].get('tensor',{}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
