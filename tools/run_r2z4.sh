set -x
timeout 600 python bench.py --no-traffic --no-extras --no-cpu --opt prof_coarse=1 > gpurun_out/r2z4_c2.json 2> gpurun_out/r2z4_c2.err; tail -2 gpurun_out/r2z4_c2.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2z4_c2.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['roofline']['step_breakdown_ms'])
PY
