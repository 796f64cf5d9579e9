set -x
timeout 1200 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2z5_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z5_pytest_gpu.log
timeout 600 python bench.py --no-traffic --no-extras --no-cpu > gpurun_out/r2z5_c2.json 2> gpurun_out/r2z5_c2.err; tail -2 gpurun_out/r2z5_c2.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2z5_c2.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'], j['roofline']['frac'], j['roofline']['step_breakdown_ms'], j['config'].get('recall_at_10'))
PY
timeout 600 python bench.py --no-traffic --no-extras --no-cpu --opt prof_coarse=1 > gpurun_out/r2z5_c2b.json 2> gpurun_out/r2z5_c2b.err; tail -2 gpurun_out/r2z5_c2b.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2z5_c2b.json') if l.startswith('{')][-1])
print('coarse stages', j['value'], j['ms_per_step'], j['roofline']['step_breakdown_ms'])
PY
