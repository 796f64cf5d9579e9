set -x
python -m pytest tests -m gpu -x -q -s -k "probe_pruning" > gpurun_out/r01j_pytest_prune.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r01j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01j_pytest_gpu.log
python bench.py > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err; tail -c 600 gpurun_out/r01j_bench.err
python bench.py --no-cpu --opt fast_prune=0 > gpurun_out/r01j_bench_noprune.json 2>> gpurun_out/r01j_bench.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --ncu-region > gpurun_out/r01j_ncu_launches.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_pass_kernel -c 3 -o gpurun_out/r01j_tc_pass python bench.py --steps 1 --warmup 3 --no-cpu --ncu-region > gpurun_out/r01j_ncu_full.log 2>&1
ncu -i gpurun_out/r01j_tc_pass.ncu-rep --page raw --csv > gpurun_out/r01j_tc_pass_raw.csv 2>/dev/null
head -c 1500 gpurun_out/r01j_bench.json
