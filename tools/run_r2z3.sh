set -x
timeout 1200 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2z3_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z3_pytest_gpu.log
timeout 600 python bench.py --no-traffic --no-extras --no-cpu > gpurun_out/r2z3_c2.json 2> gpurun_out/r2z3_c2.err; tail -2 gpurun_out/r2z3_c2.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2z3_c2.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'], j['roofline']['frac'], j['roofline']['step_breakdown_ms'], j['config'].get('recall_at_10'))
print({k:v for k,v in j['roofline'].items() if 'units' in k or 'items' in k})
PY
timeout 500 python tools/bench_small.py > gpurun_out/r2z3_small.json 2> gpurun_out/r2z3_small.err; tail -2 gpurun_out/r2z3_small.err
python - <<'PY'
import json
for line in open('gpurun_out/r2z3_small.json').read().strip().splitlines():
    try:
        l=json.loads(line); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in l.items() if k in ('index','mode','queries_per_call','median_us','equals_large_batch_exact','exact_fallbacks_per_call')} )
    except Exception as e: print(line[:200])
PY
