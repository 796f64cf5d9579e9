set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 210000 -c 35 --csv --log-file gpurun_out/r2s_kpp_launches.csv python tools/probe_kpp.py > gpurun_out/r2s_kpp.json 2> gpurun_out/r2s_kpp.err; tail -3 gpurun_out/r2s_kpp.err
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2s_kpp_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
