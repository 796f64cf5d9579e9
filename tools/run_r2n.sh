set -x
timeout 400 python tools/bench_small.py --indexes flat1m,flat1mc --modes fast --batches 9,16,32,64 > gpurun_out/r2n_small.json 2> gpurun_out/r2n_small.err; tail -2 gpurun_out/r2n_small.err
python - <<'PY'
import json
for line in open('gpurun_out/r2n_small.json').read().strip().splitlines():
    try:
        l=json.loads(line); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in l.items() if k in ('index','mode','queries_per_call','median_us','equals_large_batch_exact','fast_served_per_call','exact_fallbacks_per_call')} )
    except Exception as e: print(line[:200])
PY
