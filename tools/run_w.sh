set -x
python bench.py > gpurun_out/r01w_bench.json 2> gpurun_out/r01w_bench.err; tail -c 300 gpurun_out/r01w_bench.err
python -m pytest tests/test_gpu_fast.py tests/test_gpu_small_batch.py -x -q 2>&1 | tail -2
python - <<'P'
import json
j=json.loads(open('gpurun_out/r01w_bench.json').read().strip().splitlines()[-1])
r=j['roofline']; print(round(j['value']), round(j['e2e']['value']), j['ms_per_step'], r['bound'], round(r['frac'],3), r['algorithmic_bytes_per_launch'], r['traffic'], r['units_per_launch'], r['m64_units_per_launch'], r['why_this_bound'])
P
