#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for t in default 512,4,8 512,4,12 1024,8,6 256,2,16 384,3,8 512,4,4 1024,8,9 128,1,16; do
  timeout 120 python tools/tiny_tune.py $t 2>&1 | tail -1
done; done | tee $OUT/r02ac_tiny_tune.txt
