set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2i_pytest_gpu.log
timeout 500 python bench.py --steps 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 600 gpurun_out/r2i_bench.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
    r=l['roofline']
    print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'launches',l['gpu_launches'],'recall',l['config']['recall_at_10'],'build_s',l['config']['build_s'])
    print('frac',r['frac'],'launch_ms',r['launch_ms'],'traffic',r['traffic'],r.get('traffic_source'))
    print(r['step_breakdown_ms']); print(l['fast_vs_exact']); print('parity',l['parity']); print('cpu',l['cpu_baseline'])
    print('unpruned',l['unpruned']); print('hard',l['hard_distribution'])
except Exception as e: print('ERR',e)
PY
