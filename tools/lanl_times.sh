#!/bin/bash
# per-kernel device times of the LANL config (scale 0.25), sums in k_window / in k_sums
OUT=gpurun_out; TAG=${1:-x}
for o in 0 1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
  --log-file $OUT/lanl_launches_${TAG}_$o.csv python tools/quick_run.py config4_lanl 0.25 split_sums=$o > $OUT/ncu_lanl_$TAG.log 2>&1
echo "== split_sums=$o rc=$?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/lanl_launches_${TAG}_$o.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in d.items():
    v2=sorted(v); print(k[:40], len(v), "median %.1f us"%(v2[len(v2)//2]/1e3), "max %.1f"%(v2[-1]/1e3), "sum %.2f ms"%(sum(v)/1e6))
PY
done
