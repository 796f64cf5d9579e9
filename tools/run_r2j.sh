set -x
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --no-replicas > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; tail -c 800 gpurun_out/r2j_bench_2gpu.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2j_bench_2gpu.json').read().strip().splitlines()[-1])
    print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'recall',l['config']['recall_at_10'], l['config']['collective']['kind'])
    print(l['roofline']['frac'], l['roofline']['step_breakdown_ms']); print(l['fast_vs_exact'], l['parity']); print(l['config']['build'])
except Exception as e: print('ERR',e)
PY
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-traffic --ncu-region > gpurun_out/r2j_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2j_launches.csv 2 > gpurun_out/r2j_launches_summary.txt 2>&1; head -60 gpurun_out/r2j_launches_summary.txt
