set -x
timeout 300 python -m pytest tests/test_gpu_kpp.py -x -q > gpurun_out/r2v_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2v_pytest_gpu.log
timeout 400 python tools/probe_kpp.py > gpurun_out/r2v_kpp.json 2> gpurun_out/r2v_kpp.err; tail -3 gpurun_out/r2v_kpp.err; cat gpurun_out/r2v_kpp.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 210000 -c 7 --csv --log-file gpurun_out/r2v_kpp_launches.csv python tools/probe_kpp.py > /dev/null 2> gpurun_out/r2v_kpp2.err; tail -3 gpurun_out/r2v_kpp2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 100 -c 40 --csv --log-file gpurun_out/r2v_slice_launches.csv python tools/probe_slice.py --only "levelled x16" --reps 3 > /dev/null 2> gpurun_out/r2v_slice.err; tail -3 gpurun_out/r2v_slice.err
python - <<'PY'
import csv
for f in ('gpurun_out/r2v_kpp_launches.csv','gpurun_out/r2v_slice_launches.csv'):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    for r in rows[1:]:
        print(r[ki][:70], r[vi])
PY
