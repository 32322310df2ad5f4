set -x
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus 8 --steps 300 --warmup 5 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo "rc=$?"
MRAG_PIPELINE=0 timeout 400 $TR bench.py --gpus 8 --steps 300 --warmup 5 --no-parity > gpurun_out/r2t_bench_n8_nopipe.json 2> gpurun_out/r2t_bench_n8_nopipe.err; echo "rc=$?"
timeout 500 $TR bench.py --gpus 8 --workload c3 --steps 100 --warmup 5 > gpurun_out/r2t_bench_c3_n8_b64.json 2> gpurun_out/r2t_bench_c3_n8_b64.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items() if v}, d['gpu_launches'], (d.get('parity') or {}).get('status'), round(d['e2e']['value'],1), d.get('shard_phases_ms'), d['result_digest'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2t_bench_n8.err
