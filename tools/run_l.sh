set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01l_pytest_gpu.log 2>&1; tail -4 gpurun_out/r01l_pytest_gpu.log
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:smallscan -c 8 -o gpurun_out/r01l_smallscan python tools/bench_small.py --batches 1,8 --reps 3 --ncu-region > gpurun_out/r01l_ncu.log 2>&1
ncu -i gpurun_out/r01l_smallscan.ncu-rep --page raw --csv > gpurun_out/r01l_smallscan_raw.csv 2>/dev/null
ncu -i gpurun_out/r01l_smallscan.ncu-rep --page source --csv > gpurun_out/r01l_smallscan_src.csv 2>/dev/null
ls -la gpurun_out | tail -5
