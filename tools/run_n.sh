set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01n_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01n_pytest_gpu.log
python tools/bench_small.py > gpurun_out/r01n_small.json 2> gpurun_out/r01n_small.err; tail -3 gpurun_out/r01n_small.err
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rowstream -c 1 -o gpurun_out/r01n_rowstream python tools/bench_small.py --indexes flat1m --modes exact --batches 1 --reps 3 --ncu-region > gpurun_out/r01n_ncu.log 2>&1
ncu -i gpurun_out/r01n_rowstream.ncu-rep --page raw --csv > gpurun_out/r01n_rowstream_raw.csv 2>/dev/null
python tools/ncu_src.py gpurun_out/r01n_rowstream.ncu-rep 0 25 > gpurun_out/r01n_rowstream_src.txt 2>&1
python bench.py > gpurun_out/r01n_bench.json 2> gpurun_out/r01n_bench.err; tail -c 300 gpurun_out/r01n_bench.err
python bench.py --impl reference > gpurun_out/r01n_bench_ref.json 2>> gpurun_out/r01n_bench.err
python - <<'P'
import json
for l in open('gpurun_out/r01n_small.json'):
    j=json.loads(l); print(j['index'][:28], j['mode'], j['queries_per_call'], round(j['median_us']), 'scan', round(j['scan_us_per_call']), 'sel', round(j['select_us_per_call']), 'frac', j['scan_hbm_frac'] and round(j['scan_hbm_frac'],3), j['equals_large_batch_exact'])
P
ls -la gpurun_out
