#!/bin/bash
# A/B of k_emit build variants (libsimian_b200_<name>.so) with the split window on
OUT=gpurun_out; TAG=${1:-x}
{
echo "== default lib split_handlers=0"; timeout 300 python tools/ab_run.py config2_phold_1m 5 split_handlers=0
echo "== default lib split_handlers=1"; timeout 300 python tools/ab_run.py config2_phold_1m 5 split_handlers=1
for f in simian_b200/libsimian_b200_*.so; do
  v=$(basename $f .so); v=${v#libsimian_b200_}
  [ "$v" = user ] && continue
  echo "== $v split_handlers=1"
  SIMIAN_B200_LIB=$PWD/$f timeout 300 python tools/ab_run.py config2_phold_1m 5 split_handlers=1
done
} > $OUT/emit_variants_$TAG.txt 2>&1
cat $OUT/emit_variants_$TAG.txt
