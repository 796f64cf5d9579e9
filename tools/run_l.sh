set -x
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:smallscan --launch-skip 4 -c 2 -o gpurun_out/r01l_smallscan python tools/bench_small.py --batches 1,8 --reps 3 --ncu-region > gpurun_out/r01l_ncu.log 2>&1
ncu -i gpurun_out/r01l_smallscan.ncu-rep --page raw --csv > gpurun_out/r01l_smallscan_raw.csv 2>/dev/null
python tools/ncu_src.py gpurun_out/r01l_smallscan.ncu-rep 0 30 > gpurun_out/r01l_smallscan_src_nq1.txt 2>&1
python tools/ncu_src.py gpurun_out/r01l_smallscan.ncu-rep 1 30 > gpurun_out/r01l_smallscan_src_nq8.txt 2>&1
ls -la gpurun_out | tail -6
