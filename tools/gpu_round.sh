#!/bin/bash
# One GPU-box pass: GPU tests, bench (both arms), ncu launch list and one full capture.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
SHORT="python bench.py --draws 37888 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $SHORT > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu1_$TAG.log 2>&1
echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unfold|gemm|fold_i8|stat_slice|reduce' -s 12 -c 5 -f -o $OUT/prof_$TAG $SHORT > $OUT/ncu2_$TAG.log 2>&1
echo "ncu2 rc=$?"
cat $OUT/bench_$TAG.json | cut -c1-600
