set -x
timeout 1200 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_kpp.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2w_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2w_pytest_gpu.log
timeout 300 python tools/probe_slice.py > gpurun_out/r2w_probe.json 2> gpurun_out/r2w_probe.err; tail -3 gpurun_out/r2w_probe.err
python - <<'PY'
import json
for f in ('gpurun_out/r2w_probe.json',):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(j['rows'], j['nq'], j['k'])
    for k,v in j.items():
        if isinstance(v,dict): print(' ', k, {a:round(b,3) for a,b in v.items()})
PY
timeout 400 python tools/probe_kpp.py > gpurun_out/r2w_kpp.json 2> gpurun_out/r2w_kpp.err; tail -3 gpurun_out/r2w_kpp.err; cat gpurun_out/r2w_kpp.json
