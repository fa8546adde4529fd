#!/bin/bash
# config 4 (LANL) with every variant library, sums inside the window kernel / in k_sums
OUT=gpurun_out; TAG=${1:-x}
{
echo "== default split_sums=0"; timeout 300 python tools/quick_run.py config4_lanl 1 split_sums=0
echo "== default split_sums=1"; timeout 300 python tools/quick_run.py config4_lanl 1 split_sums=1
for f in simian_b200/libsimian_b200_*.so; do
  v=$(basename $f .so); v=${v#libsimian_b200_}
  [ "$v" = user ] && continue
  echo "== $v split_sums=1"; SIMIAN_B200_LIB=$PWD/$f timeout 300 python tools/quick_run.py config4_lanl 1 split_sums=1
done
} > $OUT/lanl_variants_$TAG.txt 2>&1
cat $OUT/lanl_variants_$TAG.txt
