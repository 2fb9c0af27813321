#!/bin/bash
# launch list (per-kernel durations) of the default bench command at B = 4,194,304
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_c5_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > gpurun_out/r02_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02_c5_launches.csv")) if len(r)>5 and r[0].isdigit()]
names=[r[4].split('<')[0].split('(')[0].replace('void ','').replace('nm::','').replace('wd::','') for r in rows]
dur=[float(r[-1].replace(',','')) for r in rows]
# one complete step = from the 2nd nm_wide_fwd_kernel to the 3rd
idx=[i for i,n in enumerate(names) if n=='nm_wide_fwd_kernel']
a,b=idx[1],idx[2]
agg=collections.OrderedDict()
for n,d in zip(names[a:b],dur[a:b]):
    k=n if n.startswith('nm_') else 'memset/other'
    c=agg.setdefault(k,[0,0.0]); c[0]+=1; c[1]+=d
tot=sum(v[1] for v in agg.values())
print("one step: %d launches, %.1f ms of kernel time" % (b-a, tot/1e6))
for k,(n,t) in agg.items(): print("  %-32s x %3d %10.2f ms %5.1f%%" % (k,n,t/1e6,100*t/tot))
PY
