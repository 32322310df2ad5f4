set -x
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2u_pytest_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest_sharded.log
tail -5 gpurun_out/r2u_pytest_sharded.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --rows 2500000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2; do
MRAG_PIPELINE=0 eval timeout 600 $T > gpurun_out/r2u_n2_nopipe_$rep.json 2> gpurun_out/r2u_err.log
MRAG_DEFER_EXCHANGE=0 eval timeout 600 $T > gpurun_out/r2u_n2_pipe_nodefer_$rep.json 2>> gpurun_out/r2u_err.log
eval timeout 600 $T > gpurun_out/r2u_n2_pipe_defer_$rep.json 2>> gpurun_out/r2u_err.log
MRAG_DEFER_EXCHANGE=0 MRAG_BENCH_PRIO=1 eval timeout 600 $T > gpurun_out/r2u_n2_pipe_nodefer_prio_$rep.json 2>> gpurun_out/r2u_err.log
MRAG_BENCH_PRIO=1 eval timeout 600 $T > gpurun_out/r2u_n2_pipe_defer_prio_$rep.json 2>> gpurun_out/r2u_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2u_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items() if v}, d['gpu_launches'], round(d['e2e']['value'],1), d['result_digest'])
    except Exception as e: print(f,'ERR',e)
PY
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2u_err.log | tail -5
