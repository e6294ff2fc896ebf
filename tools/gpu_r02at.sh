#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
bash tools/ab_lib.sh "ab_libs/lib_old.so elisa_b200/lib/libelisa_b200.so" 2 | tee $OUT/r02at_ab_f2i.txt
python bench.py --steps 2 --warmup 3 --draws 200000 --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(d['parity'])" | tee -a $OUT/r02at_ab_f2i.txt
