set -x
timeout 900 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_ivf.py -x -q > gpurun_out/r2q_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2q_pytest_gpu.log
timeout 300 python tools/probe_slice.py > gpurun_out/r2q_probe.json 2> gpurun_out/r2q_probe.err; tail -3 gpurun_out/r2q_probe.err; cat gpurun_out/r2q_probe.json
timeout 300 python tools/probe_slice.py --rows 2048 > gpurun_out/r2q_probe2k.json 2>> gpurun_out/r2q_probe.err; cat gpurun_out/r2q_probe2k.json
timeout 600 python bench.py --no-traffic > gpurun_out/r2q_c2.json 2> gpurun_out/r2q_c2.err; tail -2 gpurun_out/r2q_c2.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2q_c2.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'], j['roofline']['frac'], j['roofline']['step_breakdown_ms'])
PY
