#!/bin/bash
# round 2, first GPU pass: new tests, slice-count accuracy/time table, power per kernel class, bench
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > $OUT/r02a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02a_pytest.log
tail -5 $OUT/r02a_pytest.log
timeout 600 python tools/slices_probe.py > $OUT/r02a_slices.md 2> $OUT/r02a_slices.err; echo "slices rc=$?"
timeout 300 python tools/power_probe.py > $OUT/r02a_power.md 2> $OUT/r02a_power.err; echo "power rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02a_bench.json 2> $OUT/r02a_bench.err; echo "bench rc=$?"
ELISA_B200_I8_DEBUG=128 timeout 300 python bench.py --draws 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs > $OUT/r02a_bench_stream.json 2> $OUT/r02a_bench_stream.err; echo "stream rc=$?"
timeout 300 python bench.py --draws 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs > $OUT/r02a_bench_plain.json 2> $OUT/r02a_bench_plain.err; echo "plain rc=$?"
cut -c1-1500 $OUT/r02a_bench.json
