set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01y_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01y_pytest_gpu.log
python bench.py > gpurun_out/r01y_bench.json 2> gpurun_out/r01y_bench.err; tail -c 300 gpurun_out/r01y_bench.err
python - <<'P'
import json
j=json.loads(open('gpurun_out/r01y_bench.json').read().strip().splitlines()[-1])
r=j['roofline']; print(round(j['value']), round(j['e2e']['value']), j['ms_per_step'], r['bound'], round(r['frac'],3), j['parity'], j['fast_vs_exact']['ids_equal'], j['fast_vs_exact']['dist_bits_equal'], r['step_breakdown_ms'])
P
