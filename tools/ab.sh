#!/bin/bash
# A/B round on the GPU box: tests on the default library, then throughput of the
# named variant libraries (simian_b200/libsimian_b200_<name>.so) on the given configs.
# usage: bash tools/ab.sh <tag> "<variants>" "<configs>"
TAG=$1; VARS=$2; CFGS=${3:-config2_phold_1m}
OUT=gpurun_out; mkdir -p $OUT
timeout 420 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
for cfg in $CFGS; do
  echo "default: $(timeout 200 python tools/quick_run.py $cfg 1 2>&1 | tail -1)"
  for v in $VARS; do
    echo "$v: $(SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 200 python tools/quick_run.py $cfg 1 2>&1 | tail -1)"
  done
done 2>&1 | tee $OUT/ab_$TAG.txt
if [ -f simian_b200/libsimian_b200_prof.so ]; then
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prof.so timeout 200 python tools/phase_profile.py > $OUT/phases_$TAG.txt 2>&1
  cat $OUT/phases_$TAG.txt
fi
if [ -f simian_b200/libsimian_b200_prof1.so ]; then
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prof1.so timeout 200 python tools/burst_profile.py > $OUT/burst_$TAG.txt 2>&1
  cat $OUT/burst_$TAG.txt
fi
