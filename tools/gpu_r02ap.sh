#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python tools/tree_check.py 2>&1 | tail -3 | tee $OUT/r02ap_tree_check.txt
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
for m in edges generic; do ELISA_B200_CONTRACT=$m timeout 200 python tools/tree_rate.py 2>&1 | tail -1; done | tee $OUT/r02ap_tree_rate.txt
