#!/usr/bin/env bash
# durations of the distance-pass launches of a short run (histogram narrowing passes at the start, collecting passes after)
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dist2 -c 40 --csv --log-file gpurun_out/hist_passes.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/hist_passes.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/hist_passes.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
print(" ".join("%s:%.2f" % (r[kn].split('<')[1].split('>')[0].replace(' ',''), float(r[mv].replace(',',''))/1e6) for r in rows[hi+1:] if len(r)>mv))
PY
