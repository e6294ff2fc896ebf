#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_autograd.py -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do
for t in default 512,4,12; do
  timeout 120 python tools/tiny_tune.py $t 2>&1 | tail -1
done; done | tee $OUT/r02ad_tiny_tune.txt
python tools/small_batch_latency.py 2>&1 | tail -4 | tee $OUT/r02ad_small_batch.txt
