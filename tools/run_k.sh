mkdir -p gpurun_out/r2
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2/launches_1m.csv python bench.py --workload cfg3 --rows 1000000 --steps 3 --no-sub --no-e2e --no-cpu-baseline > gpurun_out/r2/launches_1m.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r2/launches_1m.csv') if not l.startswith('==')))
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[1:]:
    if len(r)>ix['Metric Value']: print(r[ix['Kernel Name']][:60], r[ix['Grid Size']] if 'Grid Size' in ix else '', r[ix['Metric Value']], r[ix['Metric Unit']])
PY
