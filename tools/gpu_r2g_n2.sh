set -x
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2g_pytest_2gpu_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_2gpu_sharded.log
tail -3 gpurun_out/r2g_pytest_2gpu_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_n2.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], (d.get('parity') or {}).get('status'), round(d['e2e']['value'],1), d.get('shard_phases_ms'), d['result_digest'], d['config'].get('pipeline','')[:40])
PY
