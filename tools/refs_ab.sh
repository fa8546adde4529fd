#!/bin/bash
# slots per CTA (SIMIAN_REFS variants) x entities per block, config 2
OUT=gpurun_out; TAG=${1:-x}
{
echo "== default lib"; timeout 300 python tools/ab_run.py config2_phold_1m 5
for f in simian_b200/libsimian_b200_r*.so; do
  v=$(basename $f .so); v=${v#libsimian_b200_}
  for be in 512 576 608 640; do
    echo "== $v block_entities=$be"; SIMIAN_B200_LIB=$PWD/$f timeout 300 python tools/ab_run.py config2_phold_1m 5 block_entities=$be
  done
done
} > $OUT/refs_ab_$TAG.txt 2>&1
cat $OUT/refs_ab_$TAG.txt
