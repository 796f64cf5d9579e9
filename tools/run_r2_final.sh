# The validation recipe behind profiles/r02z_* (one B200 box; the HNSW / LSH legs — tools/bench_configs.py c5, lsh — were run once
# on the earlier commit a6571f4.. and are unchanged code):  gpurun --timeout 2400 -- 'bash tools/run_r2_final.sh'
set -x
python __graft_entry__.py smoke > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r02z_pytest_gpu.log 2>&1; tail -12 gpurun_out/r02z_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02z_bench_c2.json 2> gpurun_out/r02z_bench.err; tail -c 300 gpurun_out/r02z_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02z_bench_reference_arm.json 2>> gpurun_out/r02z_bench.err
# launch list of the timed region (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02z_launches_c2_timed_region.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-traffic --ncu-region > gpurun_out/r02z_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02z_launches_c2_timed_region.csv "bench.py --steps 2 (configs[1], FAST mode): launches of the timed region" > gpurun_out/r02z_launches_c2_summary.txt 2>&1
# ncu --set full of the narrow candidate passes of one step (list sample / list scan)
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_narrow_kernel -c 2 -o gpurun_out/r02z_tc_narrow python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --no-traffic --ncu-region > gpurun_out/r02z_ncu_full.log 2>&1
ncu -i gpurun_out/r02z_tc_narrow.ncu-rep --page raw --csv > gpurun_out/r02z_tc_narrow_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/r02z_tc_narrow.ncu-rep
timeout 400 python tools/bench_small.py > gpurun_out/r02z_small_batch.json 2> gpurun_out/r02z_small.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02z_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02z_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"
timeout 300 python tools/bench_configs.py c1 > gpurun_out/r02z_config0_flat_31k.json 2> gpurun_out/r02z_cfg.err
timeout 500 python tools/bench_configs.py c3 > gpurun_out/r02z_config2_flat_10M_bf16_top100.json 2>> gpurun_out/r02z_cfg.err
tail -3 gpurun_out/r02z_cfg.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02z_*.json')):
    try:
        for line in open(f).read().strip().splitlines()[:2]:
            l=json.loads(line)
            print(f.split('/')[-1], {k:(round(v,4) if isinstance(v,float) else v) for k,v in l.items() if k in ('value','ms_per_step','value_fast','value_exact','recall_at_10','gpu_launches','fast_equals_exact','ms_per_batch')}, (l.get('roofline') or {}).get('frac'), l.get('parity') or l.get('parity_vs_oracle'))
    except Exception as e: print(f,'ERR',e)
PY
