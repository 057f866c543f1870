#!/bin/bash
# ncu launch list of a short bench run (cold-cache, serialised per-launch times)
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sharded > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -3 gpurun_out/ncu_bench_$TAG.log | cut -c1-300
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_$TAG.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        try: agg.setdefault((r[hdr.index('Kernel Name')][:44], r[hdr.index('Grid Size')]),[]).append(float(r[hdr.index('Metric Value')]))
        except Exception: pass
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1),'us')
PY
