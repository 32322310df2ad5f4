set -x
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2o_pytest_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_pytest_sharded.log
tail -5 gpurun_out/r2o_pytest_sharded.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --rows 2500000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300"
for rep in 1 2; do
for p in 0 1; do
MRAG_PIPELINE=$p eval timeout 600 $T > gpurun_out/r2o_n2_pipe${p}_$rep.json 2> gpurun_out/r2o_n2_pipe${p}_$rep.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2o_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['e2e']['value'],1), d.get('shard_phases_ms'), d['result_digest'])
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/r2o_n2_pipe1_1.err
