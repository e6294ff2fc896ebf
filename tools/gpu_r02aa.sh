#!/bin/bash
# does capping the fold kernel's registers (room for FP64-pipe blocks of the other chunk next to the persistent
# fold CTA) make two chunks in flight pay?  one stream vs two streams per library variant
set -u
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for l in elisa_b200/lib/libelisa_b200.so ab_libs/lib_r80.so ab_libs/lib_r72.so ab_libs/lib_r64.so; do
  echo -n "$l: "; ELISA_B200_LIB=$PWD/$l timeout 200 python tools/two_stream_probe.py 2>&1 | tail -1
done; done | tee $OUT/r02aa_two_streams.txt
