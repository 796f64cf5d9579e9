set -x
for noise in 0.1 0.5 1.0; do
timeout 300 python tools/bench_configs.py c5 --n 200000 --data clustered --noise $noise > gpurun_out/r2k_c5_200k_noise$noise.json 2> gpurun_out/r2k_c5.err; python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2k_c5_200k_noise$noise.json').read().strip().splitlines()[0]); print('noise $noise recall',l['recall_at_10'],'qps',l['value'],l['parity'])
except Exception as e: print('ERR',e)
PY
done
timeout 300 python tools/bench_configs.py c5 --n 200000 --data gaussian > gpurun_out/r2k_c5_200k_gauss.json 2>> gpurun_out/r2k_c5.err; python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2k_c5_200k_gauss.json').read().strip().splitlines()[0]); print('gauss recall',l['recall_at_10'],'qps',l['value'],l['parity'])
except Exception as e: print('ERR',e)
PY
tail -3 gpurun_out/r2k_c5.err
