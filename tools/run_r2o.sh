set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2o_pytest_gpu.log
timeout 400 python tools/bench_small.py --indexes flat31k,flat1m,flat1mc --modes fast --batches 9,16,32,64 > gpurun_out/r2o_small.json 2> gpurun_out/r2o_small.err; tail -2 gpurun_out/r2o_small.err
python - <<'PY'
import json
for line in open('gpurun_out/r2o_small.json').read().strip().splitlines():
    try:
        l=json.loads(line); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in l.items() if k in ('index','queries_per_call','median_us','equals_large_batch_exact','exact_fallbacks_per_call')} )
    except Exception as e: print(line[:200])
PY
timeout 300 python tools/bench_configs.py c1 > gpurun_out/r2o_c1.json 2> gpurun_out/r2o_cfg.err; python -c "
import json; l=json.loads(open('gpurun_out/r2o_c1.json').read().strip().splitlines()[0]); print('c1', l['value_fast'], l['ms_fast'], l['fast_equals_exact'], l['parity_vs_oracle'])"
timeout 300 python bench.py --steps 5 --no-cpu --no-extras --no-traffic > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; python -c "
import json; l=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1]); print('c2', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['step_breakdown_ms'], l['fast_vs_exact'])"
