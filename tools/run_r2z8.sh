set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2z8_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z8_pytest_gpu.log
timeout 600 python bench.py --no-traffic --no-extras --no-cpu > gpurun_out/r2z8_c2.json 2> gpurun_out/r2z8_c2.err; tail -2 gpurun_out/r2z8_c2.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2z8_c2.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'], j['roofline']['frac'], j['roofline']['step_breakdown_ms'], j['config'].get('recall_at_10'))
PY
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2z8_memcheck.log 2>&1; echo "memcheck rc=$?"
