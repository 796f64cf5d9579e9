set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01p_pytest_gpu.log 2>&1; tail -6 gpurun_out/r01p_pytest_gpu.log
python tools/bench_small.py > gpurun_out/r01p_small.json 2> gpurun_out/r01p_small.err; tail -3 gpurun_out/r01p_small.err
python - <<'P'
import json
for l in open('gpurun_out/r01p_small.json'):
    j=json.loads(l); print(j['index'][:28], j['mode'], j['queries_per_call'], round(j['median_us']), 'scan', round(j['scan_us_per_call']), 'sel', round(j['select_us_per_call']), 'frac', j['scan_hbm_frac'] and round(j['scan_hbm_frac'],3), j['equals_large_batch_exact'])
P
