set -x
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --opt fast_debug=1 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
grep -E "narrow pass|last tc pass" gpurun_out/r2f_bench.err | tail -8
