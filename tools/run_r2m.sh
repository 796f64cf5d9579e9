set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/r2m_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2m_pytest_gpu.log
timeout 400 python tools/bench_small.py --indexes flat1m,flat31k > gpurun_out/r2m_small.json 2> gpurun_out/r2m_small.err; tail -2 gpurun_out/r2m_small.err
python - <<'PY'
import json
for line in open('gpurun_out/r2m_small.json').read().strip().splitlines():
    try:
        l=json.loads(line); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in l.items() if k in ('index','mode','batch','nq','queries_per_call','median_us','matches_large_batch','equal')} )
    except Exception as e: print(line[:200])
PY
