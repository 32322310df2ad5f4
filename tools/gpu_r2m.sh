set -x
for g in 0 1 2; do
MRAG_GMAX=$g timeout 300 python tools/stats_probe.py 1250000 10 > gpurun_out/r2m_stats_g$g.log 2>&1
done
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2; do
for g in 0 1 2; do
MRAG_GMAX=$g eval timeout 300 $B > gpurun_out/r2m_shard_g${g}_$rep.json 2>/dev/null
done; done
B="python bench.py --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 30"
for rep in 1 2; do
for g in 0 1 2; do
MRAG_GMAX=$g eval timeout 300 $B > gpurun_out/r2m_10m_g${g}_$rep.json 2>/dev/null
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/r2m_stats_g*.log | grep -v "^+"
