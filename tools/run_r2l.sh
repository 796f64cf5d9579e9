set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/r2l_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2l_pytest_gpu.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --no-replicas --seeding random > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err; tail -c 600 gpurun_out/r2l_bench_2gpu.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2l_bench_2gpu.json').read().strip().splitlines()[-1])
    print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'recall',l['config']['recall_at_10'], l['config']['collective']['kind'])
    print(l['roofline']['frac'], l['roofline']['step_breakdown_ms']); print(l['fast_vs_exact'], l['parity'])
except Exception as e: print('ERR',e)
PY
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 5 --no-cpu --no-extras --no-traffic > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -c 400 gpurun_out/r2l_bench.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
    r=l['roofline']
    print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'launches',l['gpu_launches'],'recall',l['config']['recall_at_10'],'build_s',l['config']['build_s'])
    print('frac',r['frac'],'launch_ms',r['launch_ms'], r['items_per_launch'], r['row_tiles_read_per_launch'])
    print(r['step_breakdown_ms']); print(l['fast_vs_exact'])
except Exception as e: print('ERR',e)
PY
