#!/bin/bash
# usage (GPU box, N GPUs): bash tools/mgpu_ab.sh <N> "<option sets separated by ;>"
N=$1; OPTS=$2
IFS=';' read -ra OS <<< "$OPTS"
for o in "${OS[@]}"; do
  echo "opts [$o]: $(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_profile.py $o 2>&1 | tail -1)"
done | tee gpurun_out/mgpu_ab.txt
