for t in 1 3; do python bench.py --no-cpu --opt fast_sample_tiles=$t > gpurun_out/r01s_bench_st$t.json 2> gpurun_out/r01s.err; done
python - <<'P'
import json
for t in (1,3):
    j=json.loads(open(f'gpurun_out/r01s_bench_st{t}.json').read().strip().splitlines()[-1])
    r=j['roofline']
    print(t, round(j['value']), j['ms_per_step'], r['launch_ms'], j['fast_vs_exact']['ids_equal'], j['fast_vs_exact']['exact_fallbacks'], r['probe_pruning']['pruned_probe_pairs'], r['step_breakdown_ms'])
P
