set -x
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_pass_kernel -c 3 -o gpurun_out/r01v_tc_pass python bench.py --steps 1 --warmup 3 --no-cpu --ncu-region > gpurun_out/r01v_ncu_full.log 2>&1
ncu -i gpurun_out/r01v_tc_pass.ncu-rep --page raw --csv > gpurun_out/r01v_tc_pass_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
