#!/bin/bash
# per-kernel device times of config 4 at full size (ncu launch list)
OUT=gpurun_out; TAG=${1:-x}; shift
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $OUT/lanl_full_launches_$TAG.csv python tools/quick_run.py config4_lanl 1 "$@" > $OUT/ncu_lanl_full_$TAG.log 2>&1
echo "rc=$?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/lanl_full_launches_$TAG.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in d.items():
    v2=sorted(v); print(k[:40], len(v), "median %.1f us"%(v2[len(v2)//2]/1e3), "max %.1f"%(v2[-1]/1e3), "sum %.2f ms"%(sum(v)/1e6))
PY
