set -x
timeout 900 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2u_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2u_pytest_gpu.log
timeout 300 python tools/probe_slice.py > gpurun_out/r2u_probe.json 2> gpurun_out/r2u_probe.err; tail -3 gpurun_out/r2u_probe.err
python - <<'PY'
import json
for f in ('gpurun_out/r2u_probe.json',):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(j['rows'], j['nq'], j['k'])
    for k,v in j.items():
        if isinstance(v,dict): print(' ', k, {a:round(b,3) for a,b in v.items()})
PY
timeout 600 python bench.py --no-traffic > gpurun_out/r2u_c2.json 2> gpurun_out/r2u_c2.err; tail -2 gpurun_out/r2u_c2.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2u_c2.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'], j['roofline']['frac'], j['roofline']['step_breakdown_ms'], j['config'].get('recall_at_10'), j.get('parity'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 210000 -c 14 --csv --log-file gpurun_out/r2u_kpp_launches.csv python tools/probe_kpp.py > gpurun_out/r2u_kpp.json 2> gpurun_out/r2u_kpp.err; tail -3 gpurun_out/r2u_kpp.err
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2u_kpp_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
