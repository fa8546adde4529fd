#!/bin/bash
# A/B of every simian_b200/libsimian_b200_*.so against the default library (config 2)
OUT=gpurun_out; TAG=${1:-x}; REPS=${2:-2}
{
for i in $(seq $REPS); do
echo "== default"; timeout 300 python tools/ab_run.py config2_phold_1m 5
for f in simian_b200/libsimian_b200_*.so; do
  v=$(basename $f .so); v=${v#libsimian_b200_}
  [ "$v" = user ] && continue
  echo "== $v"; SIMIAN_B200_LIB=$PWD/$f timeout 300 python tools/ab_run.py config2_phold_1m 5
done
done
} > $OUT/ab_libs_$TAG.txt 2>&1
cat $OUT/ab_libs_$TAG.txt
