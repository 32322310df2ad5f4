set -x
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2; do
for l in 0 1; do
MRAG_CARVEOUT=$l eval timeout 300 $B > gpurun_out/r2r_shard_carve${l}_$rep.json 2>gpurun_out/r2r_err.log
done; done
for l in 0 1; do
MRAG_CARVEOUT=$l timeout 300 python bench.py --workload c2 --no-cpu-baseline --sweep '' --threads 0 --no-parity > gpurun_out/r2r_c2_carve${l}.json 2>>gpurun_out/r2r_err.log
done
timeout 300 python tools/stats_probe.py 1250000 10 > gpurun_out/r2r_stats.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], d['gpu_launches'], d['roofline'] and round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2r_err.log; cat gpurun_out/r2r_stats.log
