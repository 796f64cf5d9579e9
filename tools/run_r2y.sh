set -x
for hf in 0 5120 3456 2560; do
timeout 300 python bench.py --no-traffic --no-extras --no-cpu --opt host_feed=$hf > gpurun_out/r2y_c2_hf$hf.json 2> gpurun_out/r2y_c2_hf$hf.err; tail -2 gpurun_out/r2y_c2_hf$hf.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/r2y_c2_hf$hf.json') if l.startswith('{')][-1])
print('host_feed', $hf, j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], j['gpu_launches'], j.get('parity'))
PY
done
