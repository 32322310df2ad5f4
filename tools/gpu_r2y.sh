set -x
timeout 900 python -m pytest tests/test_hybrid.py tests/test_gpu_corpus_search.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2y_pytest_hybrid.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_pytest_hybrid.log
tail -5 gpurun_out/r2y_pytest_hybrid.log
timeout 600 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2y_c5_pairs.json 2> gpurun_out/r2y_c5_pairs.err
MRAG_HYB_PAIRS=0 timeout 600 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2y_c5_scan.json 2> gpurun_out/r2y_c5_scan.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y_c5*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d['phases_ms'], d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['config']['mean_rows_returned'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2y_c5_pairs.err
