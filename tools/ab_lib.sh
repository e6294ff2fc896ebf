#!/bin/bash
# A/B of library builds on ONE box (kernel variants compiled with different -D flags into ab_libs/):
# usage: bash tools/ab_lib.sh "lib1.so lib2.so ..." [reps] [extra bench args]; prints per-class ms per step.
LIBS=$1; REPS=${2:-2}; shift 2
for r in $(seq 1 $REPS); do
  for l in $LIBS; do
    ELISA_B200_LIB=$PWD/$l python bench.py --draws 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); c=d['roofline']['classes_ms_per_step']
print('$l', 'evals/s %.4g' % d['value'], ' '.join('%s %.2f' % (k, x) for k, x in c.items()), 'MHz', d['clocks']['sm_mhz'])"
  done
done
