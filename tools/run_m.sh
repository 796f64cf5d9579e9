set -x
python tools/bench_small.py --indexes flat1m --modes exact --batches 1,8 --reps 20 --grid "stream_seg=256;stream_warps=8,10,12" > gpurun_out/r01m_sweep2.json 2> gpurun_out/r01m_sweep2.err; tail -3 gpurun_out/r01m_sweep2.err
python tools/bench_small.py --indexes flat1m --modes exact --batches 1,8 --reps 20 --grid "stream_seg=384;stream_warps=6,8" >> gpurun_out/r01m_sweep2.json 2>> gpurun_out/r01m_sweep2.err
python tools/bench_small.py --indexes flat1m --modes exact --batches 1,8 --reps 20 --grid "stream_seg=512;stream_warps=4,5,6" >> gpurun_out/r01m_sweep2.json 2>> gpurun_out/r01m_sweep2.err
python - <<'P'
import json
for l in open('gpurun_out/r01m_sweep2.json'):
    j=json.loads(l); print(j['knobs'], j['queries_per_call'], round(j['median_us']), 'scan', round(j['scan_us_per_call']), 'sel', round(j['select_us_per_call']), 'frac', j['scan_hbm_frac'] and round(j['scan_hbm_frac'],3), j['equals_large_batch_exact'])
P
