set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ksplit or default_dispatch or scan_paths or sampling" > gpurun_out/r2c_pytest_ks.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_ks.log
tail -3 gpurun_out/r2c_pytest_ks.log
timeout 600 python bench.py --rows 6250000 --dim 1536 --batch 64 --sweep '1,16' --threads 0 --steps 20 --no-parity > gpurun_out/r2c_bench_1536_b64.json 2> gpurun_out/r2c_bench_1536_b64.err; echo "rc=$?"
for v in default noqready; do
  if [ $v = noqready ]; then export MRAG_LIB=$PWD/build_ab/libmrag_noqready.so; fi
  timeout 300 python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 200 --no-parity > gpurun_out/r2c_shard1250k_$v.json 2> gpurun_out/r2c_shard1250k_$v.err
  timeout 300 python bench.py --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 40 --no-parity > gpurun_out/r2c_n1_$v.json 2> gpurun_out/r2c_n1_$v.err
done
unset MRAG_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],3), [(x['batch'],round(x['qps']),round(x['scan_frac_of_hbm_peak'],3)) for x in d['sweep']])
    except Exception as e: print(f,'ERR',e)
PY
