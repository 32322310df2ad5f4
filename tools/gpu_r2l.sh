set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2; do
for g in 0 1 2; do
MRAG_GMAX=$g eval timeout 300 $B > gpurun_out/r2l_shard_g${g}_$rep.json 2>/dev/null
done; done
B="python bench.py --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 30"
for rep in 1 2; do
for g in 0 1 2; do
MRAG_GMAX=$g eval timeout 300 $B > gpurun_out/r2l_10m_g${g}_$rep.json 2>/dev/null
done; done
B="python bench.py --rows 6250000 --dim 1536 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 30"
for g in 0 1 2; do
MRAG_GMAX=$g eval timeout 300 $B > gpurun_out/r2l_1536_g${g}.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], d['gpu_launches'], d.get('parity',{}).get('status'))
    except Exception as e: print(f,'ERR',e)
PY
