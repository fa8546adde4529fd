#!/bin/bash
# per-kernel device times of the split window (ncu launch list; serialised, cold cache)
TAG=${1:-x}; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/launches_$TAG.csv python tools/quick_run.py config2_phold_1m 1 split_handlers=1 "$@" > $OUT/ncu1_$TAG.log 2>&1
echo "ncu launches rc=$?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches_$TAG.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in d.items():
    v2=sorted(v); print(k[:40], len(v), "median %.1f us"%(v2[len(v2)//2]/1e3), "min %.1f"%(v2[0]/1e3), "sum %.2f ms"%(sum(v)/1e6))
PY
