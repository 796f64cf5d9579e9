# N-GPU box: the multi-GPU tests, then bench.py --gpus N (sharded 12.5 M rows per GPU)
set -x
N=$1
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02z_pytest_multi.log 2>&1; tail -3 gpurun_out/r02z_pytest_multi.log; fi
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02z_bench_${N}gpu.json 2> gpurun_out/r02z_bench_${N}gpu.err; tail -c 400 gpurun_out/r02z_bench_${N}gpu.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/r02z_bench_${N}gpu.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['config'].get('recall_at_10'), j['roofline']['frac'])
print(j['roofline']['step_breakdown_ms'])
print({k:v for k,v in j['config']['build'].items() if k in ('build_s','seeding_s','assign_ms')})
print(j.get('parity'))
PY
