# round 2: the N > 1 default workload (12.5M rows per GPU) on 2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash tools/run_r2b.sh'
set -x
(while true; do nvidia-smi --query-gpu=index,memory.used --format=csv,noheader; sleep 5; done) > gpurun_out/r2b_mem.log 2>&1 &
MON=$!
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err; tail -c 1500 gpurun_out/r2b_bench_2gpu.err; tail -c 5000 gpurun_out/r2b_bench_2gpu.json
kill $MON
sort -t, -k2 -n -r gpurun_out/r2b_mem.log | head -3
