#!/bin/bash
# usage (GPU box, N GPUs): bash tools/mgpu.sh <tag> <N> [skip_tests]
TAG=$1; N=$2
OUT=gpurun_out; mkdir -p $OUT
if [ -z "$3" ]; then
  timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q > $OUT/pytest_dist_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -15 $OUT/pytest_dist_$TAG.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 2 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench rc=$?"; cat $OUT/bench_${TAG}_n$N.json; tail -5 $OUT/bench_${TAG}_n$N.err
