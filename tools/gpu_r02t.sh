#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
bash tools/ab_lib.sh "ab_libs/lib_stat11.so ab_libs/lib_new11.so ab_libs/lib_stat11_mb5.so ab_libs/lib_stat11_mb6.so" 2 > $OUT/r02t_ab_stat2.txt 2>&1
cat $OUT/r02t_ab_stat2.txt
