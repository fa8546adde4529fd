#!/bin/bash
# A/B of the split window (k_window_stage + k_emit) against the one-kernel window.
# usage (GPU box): bash tools/split_ab.sh <tag> [skip_tests]
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q --timeout 180 > $OUT/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log
fi
{
for o in "split_handlers=0" "split_handlers=1"; do
  echo "== $o"; timeout 300 python tools/ab_run.py config2_phold_1m 5 $o
done
for v in emit4 emit6; do
  if [ -f simian_b200/libsimian_b200_$v.so ]; then
    echo "== $v split_handlers=1"
    SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_$v.so timeout 300 python tools/ab_run.py config2_phold_1m 5 split_handlers=1
  fi
done
for cfg in config5_phold_4m_small_lookahead config1_phold_1k; do
  for o in "split_handlers=0" "split_handlers=1"; do
    echo "== $cfg $o"; timeout 300 python tools/ab_run.py $cfg 3 $o
  done
done
} > $OUT/split_ab_$TAG.txt 2>&1
cat $OUT/split_ab_$TAG.txt
