# round 2: 100M x 768 on 8 GPUs (and the 4-GPU point):  gpurun --gpus 8 --timeout 900 -- 'bash tools/run_r2c.sh'
set -x
(while true; do nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | sort -t, -k2 -n -r | head -1; sleep 5; done) > gpurun_out/r2c_mem.log 2>&1 &
MON=$!
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c_bench_8gpu.json 2> gpurun_out/r2c_bench_8gpu.err; tail -c 1500 gpurun_out/r2c_bench_8gpu.err; tail -c 6000 gpurun_out/r2c_bench_8gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 --no-replicas > gpurun_out/r2c_bench_4gpu.json 2> gpurun_out/r2c_bench_4gpu.err; tail -c 500 gpurun_out/r2c_bench_4gpu.err; head -c 1200 gpurun_out/r2c_bench_4gpu.json
kill $MON
sort -t, -k2 -n -r gpurun_out/r2c_mem.log | head -2
