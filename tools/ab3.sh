#!/bin/bash
# usage: bash tools/ab3.sh <tag> "<variants>" <config> <reps> "<option sets separated by ;>"
TAG=$1; VARS=$2; CFG=${3:-config2_phold_1m}; REPS=${4:-5}; OPTS=$5
OUT=gpurun_out; mkdir -p $OUT
{
timeout 200 python tools/ab_run.py $CFG $REPS 2>&1 | tail -1
IFS=';' read -ra OS <<< "$OPTS"
for o in "${OS[@]}"; do
  echo "opts $o: $(timeout 200 python tools/ab_run.py $CFG $REPS $o 2>&1 | tail -1)"
done
for v in $VARS; do
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 200 python tools/ab_run.py $CFG $REPS 2>&1 | tail -1
done
} | tee $OUT/ab3_$TAG.txt
