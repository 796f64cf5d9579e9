# The validation recipe behind profiles/r01r_* .. r01y_* (one B200 box):  gpurun --timeout 1500 -- 'bash tools/run_r.sh'
set -x
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
# launch list of the timed region (cold-cache, serialised: shares, not absolutes)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --ncu-region > gpurun_out/ncu_launches.log 2>&1
# ncu --set full of the candidate passes of one step (coarse dump / list sample / list scan)
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_pass_kernel -c 3 -o gpurun_out/tc_pass python bench.py --steps 1 --warmup 3 --no-cpu --ncu-region > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/tc_pass.ncu-rep --page raw --csv > gpurun_out/tc_pass_raw.csv 2>/dev/null
python tools/bench_small.py > gpurun_out/small.json 2> gpurun_out/small.err
timeout 300 python tools/bench_configs.py c1 > gpurun_out/c1.json 2> gpurun_out/c1.err
timeout 400 python tools/bench_configs.py c5 --n 1000000 > gpurun_out/c5.json 2> gpurun_out/c5.err
timeout 500 python tools/bench_configs.py c3 > gpurun_out/c3.json 2> gpurun_out/c3.err
timeout 300 python tools/bench_configs.py lsh > gpurun_out/lsh.json 2> gpurun_out/lsh.err
