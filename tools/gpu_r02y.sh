#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_autograd.py tests/test_gpu_nccl.py tests/test_gpu_conv.py -m gpu -q > $OUT/r02y_new.log 2>&1; echo "tests rc=$?" | tee -a $OUT/r02y_new.log
tail -n 15 $OUT/r02y_new.log
python tools/small_batch_latency.py > $OUT/r02y_small_batch.txt 2>&1; tail -n 6 $OUT/r02y_small_batch.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tiny_eval' -c 6 python tools/tiny_ncu_run.py 2>&1 | grep -i "tiny_eval\|gpu__time" | head -12
