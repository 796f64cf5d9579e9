set -x
python __graft_entry__.py smoke > gpurun_out/r01r_smoke.log 2>&1; tail -2 gpurun_out/r01r_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r01r_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01r_pytest_gpu.log
python bench.py > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err; tail -c 300 gpurun_out/r01r_bench.err
python bench.py --impl reference > gpurun_out/r01r_bench_ref.json 2>> gpurun_out/r01r_bench.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --ncu-region > gpurun_out/r01r_ncu_launches.log 2>&1
python tools/bench_small.py > gpurun_out/r01r_small.json 2> gpurun_out/r01r_small.err
timeout 300 python tools/bench_configs.py c1 > gpurun_out/r01r_c1.json 2> gpurun_out/r01r_c1.err
timeout 400 python tools/bench_configs.py c5 --n 1000000 > gpurun_out/r01r_c5.json 2> gpurun_out/r01r_c5.err
timeout 500 python tools/bench_configs.py c3 > gpurun_out/r01r_c3.json 2> gpurun_out/r01r_c3.err
tail -c 400 gpurun_out/r01r_c1.json; tail -c 500 gpurun_out/r01r_c5.json; tail -c 500 gpurun_out/r01r_c3.json
