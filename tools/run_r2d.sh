# round 2: new tests (a4 / PCAF / bf16 IVF / full-size / multi one-rank) + compute-sanitizer + SASS listing on one GPU
set -x
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -25 gpurun_out/r2d_pytest_gpu.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2d_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2d_sanitizer_memcheck_smoke.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2d_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2d_sanitizer_racecheck_smoke.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pcaf.py -x -q -k "one_rank or pcaf_projection or lane_distances" > gpurun_out/r2d_sanitizer_memcheck_multi_pcaf.log 2>&1; echo "memcheck2 rc=$?"; tail -6 gpurun_out/r2d_sanitizer_memcheck_multi_pcaf.log
cuobjdump -sass hnsw_clj_b200/build/hb_tc.o | grep -E "Function :|UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS" | awk '{a[$0]++} END {for (k in a) print a[k], k}' | sort -k2 | head -100 > gpurun_out/r2d_sass_tc_pass_summary.txt; wc -l gpurun_out/r2d_sass_tc_pass_summary.txt
