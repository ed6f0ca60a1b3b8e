#!/bin/bash
# per-kernel durations of one bench run (ncu launch list; cold-cache, serialised)
W=${1:-cfg3_B32_NH4_S1600_DH128}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/plain_k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/klist.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/ncu_k.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/klist.csv') if not l.startswith('==')]
agg=collections.defaultdict(list)
for row in csv.DictReader(lines):
    agg[row['Kernel Name'][:60]].append(float(row['Metric Value'].replace(',',''))/1e3)
for n,v in sorted(agg.items(), key=lambda x:-sum(x[1]))[:10]:
    print(f"{sum(v)/len(v):9.1f} us avg  x{len(v):3d}  {n}")
PY
