#!/bin/bash
# usage (GPU box): bash tools/tune.sh <tag> "<variant names>" [config] [extra quick_run args]
TAG=$1; VARS=$2; CFG=${3:-config2_phold_1m}; shift; shift; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
echo "default: $(timeout 300 python tools/quick_run.py $CFG 1 "$@" 2>&1 | tail -1)"
for v in $VARS; do
  echo "$v: $(SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 300 python tools/quick_run.py $CFG 1 "$@" 2>&1 | tail -1)"
done
if [ -f simian_b200/libsimian_b200_prof.so ]; then
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prof.so timeout 300 python tools/phase_profile.py $CFG > $OUT/phases_$TAG.txt 2>&1
  cat $OUT/phases_$TAG.txt
fi
