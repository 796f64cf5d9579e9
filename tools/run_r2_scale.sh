# bench.py --gpus N on an N-GPU box (profiles/r02z_bench_<N>gpu_*.json.log):  gpurun --gpus N --timeout 900 -- 'bash tools/run_r2_scale.sh N'
set -x
N=$1
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02z_bench_${N}gpu.json 2> gpurun_out/r02z_bench_${N}gpu.err; tail -c 600 gpurun_out/r02z_bench_${N}gpu.err
