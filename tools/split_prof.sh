#!/bin/bash
# per-kernel times + ncu full capture of the split window (k_window_stage, k_emit)
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/ab_run.py config2_phold_1m 3 split_handlers=0 > $OUT/split_prof_$TAG.txt 2>&1
timeout 300 python tools/ab_run.py config2_phold_1m 3 split_handlers=1 >> $OUT/split_prof_$TAG.txt 2>&1
cat $OUT/split_prof_$TAG.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python tools/quick_run.py config2_phold_1m 1 split_handlers=1 > $OUT/ncu1_$TAG.log 2>&1
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
    v2=sorted(v); print(k[:40], len(v), "median %.1f us"%(v2[len(v2)//2]/1e3), "sum %.2f ms"%(sum(v)/1e6))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_window_stage|k_emit" -s 120 -c 4 \
  -o $OUT/prof_$TAG -f python tools/quick_run.py config2_phold_1m 1 split_handlers=1 > $OUT/ncu2_$TAG.log 2>&1
echo "ncu full rc=$?"
