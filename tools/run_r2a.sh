# round 2, first multi-GPU check (2 GPUs):  gpurun --gpus 2 --timeout 1200 -- 'bash tools/run_r2a.sh'
set -x
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2a_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --shard-workload small > gpurun_out/r2a_bench_2gpu_small.json 2> gpurun_out/r2a_bench_2gpu_small.err; tail -c 1500 gpurun_out/r2a_bench_2gpu_small.err; tail -c 3000 gpurun_out/r2a_bench_2gpu_small.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --shard-workload small --no-replicas --opt comm_p2p=0 > gpurun_out/r2a_bench_2gpu_small_nccl.json 2> gpurun_out/r2a_bench_2gpu_small_nccl.err; tail -c 800 gpurun_out/r2a_bench_2gpu_small_nccl.err; tail -c 2000 gpurun_out/r2a_bench_2gpu_small_nccl.json
timeout 400 python bench.py --steps 3 > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err; tail -c 600 gpurun_out/r2a_bench_1gpu.err; head -c 600 gpurun_out/r2a_bench_1gpu.json
