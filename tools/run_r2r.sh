set -x
timeout 900 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_kpp.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2r_pytest_gpu.log
timeout 300 python tools/probe_slice.py > gpurun_out/r2r_probe.json 2> gpurun_out/r2r_probe.err; tail -3 gpurun_out/r2r_probe.err; cat gpurun_out/r2r_probe.json
timeout 400 python tools/probe_kpp.py > gpurun_out/r2r_kpp.json 2> gpurun_out/r2r_kpp.err; tail -3 gpurun_out/r2r_kpp.err; cat gpurun_out/r2r_kpp.json
