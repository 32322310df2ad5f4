set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "mma128 or pair_scan or scan_paths or ksplit or sampling or default_dispatch" > gpurun_out/r2k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
C="python bench.py --workload c3 --batch 1024 --sweep '' --threads 0 --steps 10 --no-parity --no-cpu-baseline"
for rep in 1 2; do
eval timeout 300 $C > gpurun_out/r2k_c3_b1024_kc2k_$rep.json 2>/dev/null
MRAG_KC_2K=0 eval timeout 300 $C > gpurun_out/r2k_c3_b1024_kc15k_$rep.json 2>/dev/null
done
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2 3; do
eval timeout 300 $B > gpurun_out/r2k_shard_qhl_$rep.json 2>/dev/null
MRAG_QHL=0 eval timeout 300 $B > gpurun_out/r2k_shard_noqhl_$rep.json 2>/dev/null
done
timeout 600 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2k_bench_c5.json 2> gpurun_out/r2k_bench_c5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
PY
