#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_random.py tests/test_gpu_fit.py -m gpu -q -x 2>&1 | tail -8
for rep in 1 2; do for m in edges generic; do ELISA_B200_CONTRACT=$m timeout 200 python tools/tree_rate.py 2>&1 | tail -1; done; done | tee $OUT/r02an_tree_rate.txt
