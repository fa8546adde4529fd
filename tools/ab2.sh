#!/bin/bash
# usage: bash tools/ab2.sh <tag> "<variants>" [config] [reps]
TAG=$1; VARS=$2; CFG=${3:-config2_phold_1m}; REPS=${4:-5}
OUT=gpurun_out; mkdir -p $OUT
{
timeout 200 python tools/ab_run.py $CFG $REPS 2>&1 | tail -1
for v in $VARS; do
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 200 python tools/ab_run.py $CFG $REPS 2>&1 | tail -1
done
} | tee $OUT/ab2_$TAG.txt
