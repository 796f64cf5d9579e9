# round 2: narrow-unit kernel check on one GPU
set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -12 gpurun_out/r2e_pytest_gpu.log
timeout 300 python bench.py --steps 5 --no-cpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 600 gpurun_out/r2e_bench.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
    r=l['roofline']
    print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'launches',l['gpu_launches'],'recall',l['config']['recall_at_10'])
    print('frac',r['frac'],'launch_ms',r['launch_ms'],'narrow',r.get('narrow_units_per_launch'),r.get('narrow_items_per_launch'),'items',r['items_per_launch'])
    print(r['step_breakdown_ms']); print(l['fast_vs_exact'])
except Exception as e: print('ERR',e)
PY
timeout 300 python bench.py --steps 5 --no-cpu --opt tc_narrow=0 > gpurun_out/r2e_bench_nonarrow.json 2>> gpurun_out/r2e_bench.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2e_bench_nonarrow.json').read().strip().splitlines()[-1])
    r=l['roofline']
    print('NO NARROW value',l['value'],'ms',l['ms_per_step'],'frac',r['frac'],'launch_ms',r['launch_ms']); print(r['step_breakdown_ms'])
except Exception as e: print('ERR',e)
PY
