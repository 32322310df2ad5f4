set -x
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2; do
for l in 2 1 0; do
MRAG_EVENTS=$l eval timeout 300 $B > gpurun_out/r2q_shard_ev${l}_$rep.json 2>gpurun_out/r2q_err.log
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], d['gpu_launches'], d['roofline'] and round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2q_err.log
