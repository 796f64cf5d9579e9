set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01q_pytest_gpu.log 2>&1; tail -6 gpurun_out/r01q_pytest_gpu.log
python bench.py --no-cpu > gpurun_out/r01q_bench.json 2> gpurun_out/r01q_bench.err; tail -c 400 gpurun_out/r01q_bench.err
python bench.py --no-cpu --opt tc_half_m=0 > gpurun_out/r01q_bench_m128.json 2>> gpurun_out/r01q_bench.err
python - <<'P'
import json
for f in ['gpurun_out/r01q_bench.json','gpurun_out/r01q_bench_m128.json']:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    r=j['roofline']
    print(f, round(j['value']), round(j['e2e']['value']), j['ms_per_step'], r['bound'], r['frac'], r['launch_ms'], j['fast_vs_exact'], r['step_breakdown_ms'])
P
