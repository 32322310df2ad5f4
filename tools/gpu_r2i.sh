set -x
timeout 1200 python -m pytest tests/test_gpu_corpus_search.py tests/test_gpu_pool.py -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log
tail -30 gpurun_out/r2i_pytest.log
MRAG_DEVICES=0,0 timeout 300 python tools/multi_c4.py 500000 5 > gpurun_out/r2i_multi_c4_small.json 2> gpurun_out/r2i_multi_c4_small.err; echo "rc=$?"; tail -3 gpurun_out/r2i_multi_c4_small.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sampling or mma128 or pair_scan or scan_paths or ksplit" > gpurun_out/r2i_pytest_sample4.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest_sample4.log
tail -3 gpurun_out/r2i_pytest_sample4.log
B="python bench.py --rows 1250000 --no-cpu-baseline --sweep '' --threads 0 --also-f32 0 --steps 300 --no-parity"
for rep in 1 2; do
eval timeout 300 $B > gpurun_out/r2i_shard_default_$rep.json 2>/dev/null
MRAG_LIB=$PWD/build_ab/libmrag_samplek16.so eval timeout 300 $B > gpurun_out/r2i_shard_samplek16_$rep.json 2>/dev/null
MRAG_SAMPLE_MIN_TILES=100000000 eval timeout 300 $B > gpurun_out/r2i_shard_unsampled_reg_$rep.json 2>/dev/null
MRAG_SAMPLE_MIN_TILES=100000000 MRAG_UNSAMPLED_BUFFER=1 eval timeout 300 $B > gpurun_out/r2i_shard_unsampled_buf_$rep.json 2>/dev/null
done
L="python bench.py --no-cpu-baseline --threads 0 --also-f32 0 --steps 20 --no-parity --sweep 256,1024"
eval timeout 300 $L > gpurun_out/r2i_n1_sample4.json 2>/dev/null
MRAG_LIB=$PWD/build_ab/libmrag_samplek16.so eval timeout 300 $L > gpurun_out/r2i_n1_sample16.json 2>/dev/null
eval timeout 300 $L --workload c3 > gpurun_out/r2i_c3_sample4.json 2>/dev/null
MRAG_LIB=$PWD/build_ab/libmrag_samplek16.so eval timeout 300 $L --workload c3 > gpurun_out/r2i_c3_sample16.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d: print(f, round(d['value']), round(d['ms_per_step'],4), d['phases_ms'], d['gpu_launches'], [(x['batch'], round(x['qps']), round(x['ms_per_step'],3)) for x in d.get('sweep',[])])
        else: print(f, {k: d[k] for k in d if k.endswith('_ms') or k.startswith('hbm') or k=='unfiltered_frac_of_hbm_floor'})
    except Exception as e: print(f,'ERR',e)
PY
