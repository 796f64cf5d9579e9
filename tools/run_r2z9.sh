set -x
python __graft_entry__.py smoke > gpurun_out/r2z9_smoke.log 2>&1; tail -1 gpurun_out/r2z9_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z9_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2z9_pytest_gpu.log
