#!/bin/bash
# per-kernel durations (ncu launch list, caches and clocks left alone) of the adjoint at B = 37,888 for timing-experiment variants
for v in "" nostage nostash wrap128; do
  if [ -z "$v" ]; then unset NM_LIB_PATH; else export NM_LIB_PATH=neuromancer_b200/_C/variants/lib_c5_$v.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'nm_wide' -s 20 -c 12 --csv --log-file gpurun_out/r02b_l_$v.csv \
     python bench.py --workload c5 --batch 37888 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02b_l_$v.csv")) if len(r)>5 and r[0].isdigit()]
out={}
for r in rows:
    name=r[4].split('<')[0].split('(')[0].replace('void ','').replace('nm::','')
    out.setdefault(name,[]).append(float(r[-1].replace(',','')))
print("variant [$v]", {k: round(sum(x)/len(x)/1e6,3) if max(x)>1e4 else round(sum(x)/len(x)/1e3,3) for k,x in out.items()})
PY
done
