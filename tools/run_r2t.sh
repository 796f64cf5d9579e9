set -x
timeout 300 python -m pytest tests/test_gpu_kpp.py -x -q > gpurun_out/r2t_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2t_pytest_gpu.log
timeout 300 python tools/probe_slice.py > gpurun_out/r2t_probe.json 2> gpurun_out/r2t_probe.err; tail -3 gpurun_out/r2t_probe.err
timeout 300 python tools/probe_slice.py --rows 65536 --nq 16384 --k 1 > gpurun_out/r2t_probe_assign.json 2>> gpurun_out/r2t_probe.err; tail -3 gpurun_out/r2t_probe.err
python - <<'PY'
import json
for f in ('gpurun_out/r2t_probe.json','gpurun_out/r2t_probe_assign.json'):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(j['rows'], j['nq'], j['k'])
    for k,v in j.items():
        if isinstance(v,dict): print(' ', k, {a:round(b,3) for a,b in v.items()})
PY
timeout 400 python tools/probe_kpp.py > gpurun_out/r2t_kpp.json 2> gpurun_out/r2t_kpp.err; tail -3 gpurun_out/r2t_kpp.err; cat gpurun_out/r2t_kpp.json
