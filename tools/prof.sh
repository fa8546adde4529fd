#!/bin/bash
# usage (GPU box): bash tools/prof.sh <tag> [config]   -- tests, quick run, phases, ncu full capture
TAG=$1; CFG=${2:-config2_phold_1m}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
echo "default: $(timeout 300 python tools/quick_run.py $CFG 1 2>&1 | tail -1)"
SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prof.so timeout 300 python tools/phase_profile.py $CFG > $OUT/phases_$TAG.txt 2>&1
cat $OUT/phases_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_window -s 40 -c 2 \
  -o $OUT/prof_$TAG -f python tools/quick_run.py $CFG 1 > $OUT/ncu2_$TAG.log 2>&1
echo "ncu full rc=$?"
