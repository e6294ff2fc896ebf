#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/r02c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02c_pytest.log
tail -5 $OUT/r02c_pytest.log
timeout 300 python tools/slices_probe.py --time-only > $OUT/r02c_slices_time.md 2> $OUT/r02c_slices.err; echo "slices rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02c_bench.json 2> $OUT/r02c_bench.err; echo "bench rc=$?"
cut -c1-3000 $OUT/r02c_bench.json
