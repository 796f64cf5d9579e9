set -x
timeout 900 python -m pytest tests/test_gpu_small_batch.py tests/test_gpu_fast.py tests/test_gpu_parity.py -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2p_pytest_gpu.log
timeout 500 python tools/bench_small.py > gpurun_out/r2p_small.json 2> gpurun_out/r2p_small.err; tail -2 gpurun_out/r2p_small.err
python - <<'PY'
import json
for line in open('gpurun_out/r2p_small.json').read().strip().splitlines():
    try:
        l=json.loads(line); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in l.items() if k in ('index','mode','queries_per_call','median_us','equals_large_batch_exact','exact_fallbacks_per_call')} )
    except Exception as e: print(line[:200])
PY
