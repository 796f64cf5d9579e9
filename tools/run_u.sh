set -x
python __graft_entry__.py smoke > gpurun_out/r01u_smoke.log 2>&1; tail -1 gpurun_out/r01u_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r01u_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01u_pytest_gpu.log
python bench.py > gpurun_out/r01u_bench.json 2> gpurun_out/r01u_bench.err; tail -c 300 gpurun_out/r01u_bench.err
python bench.py --impl reference > gpurun_out/r01u_bench_ref.json 2>> gpurun_out/r01u_bench.err
python - <<'P'
import json
j=json.loads(open('gpurun_out/r01u_bench.json').read().strip().splitlines()[-1])
r=j['roofline']; print(round(j['value']), round(j['e2e']['value']), j['ms_per_step'], r['bound'], round(r['frac'],3), j['cpu_baseline']['value'], j['parity'], j['fast_vs_exact'])
print(open('gpurun_out/r01u_bench_ref.json').read()[:200])
P
