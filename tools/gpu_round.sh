#!/bin/bash
# One GPU round: tests, bench, phase profile, ncu launch list + one full capture.
# usage (on the GPU box): bash tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json
if [ -f simian_b200/libsimian_b200_prof.so ]; then
  SIMIAN_B200_LIB=$PWD/simian_b200/libsimian_b200_prof.so timeout 300 python tools/phase_profile.py > $OUT/phases_$TAG.txt 2>&1
  cat $OUT/phases_$TAG.txt
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu1_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_window -s 60 -c 2 \
  -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $OUT/ncu2_$TAG.log 2>&1
echo "ncu full rc=$?"
