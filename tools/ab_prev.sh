#!/bin/bash
# A/B of the current library against simian_b200/libsimian_b200_prev.so (the previous
# commit's build) on config 2 and the LANL config
OUT=gpurun_out; TAG=${1:-x}
{
for i in 1 2; do
echo "== current"; timeout 300 python tools/ab_run.py config2_phold_1m 5
echo "== prev"; SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prev.so timeout 300 python tools/ab_run.py config2_phold_1m 5
done
echo "== current lanl"; timeout 300 python tools/quick_run.py config4_lanl 0.25
echo "== prev lanl"; SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prev.so timeout 300 python tools/quick_run.py config4_lanl 0.25
} > $OUT/ab_prev_$TAG.txt 2>&1
cat $OUT/ab_prev_$TAG.txt
