#!/bin/bash
# ncu launch list of the bench command (share of each kernel in a step; cold-cache, serialised)
TAG=${1:-r2x}
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-reps 5"
timeout 200 $CMD > $O/${TAG}_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${TAG}_ncu_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/${TAG}_ncu_launches.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=="ID"][0]
h=rows[hdr]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki][:90]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:28]:
    print(f"{a[0]:5d} x {a[1]/a[0]/1000:8.2f} us = {100*a[1]/tot:5.1f} %  {k}")
PY
