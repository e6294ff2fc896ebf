#!/bin/bash
# A/B of an environment switch on ONE box: alternates the two settings, prints per-class ms.
# usage: bash tools/ab_env.sh VAR A B [reps] [extra bench args]
VAR=$1; A=$2; B=$3; REPS=${4:-3}; shift 4
for r in $(seq 1 $REPS); do
  for v in $A $B; do
    env $VAR=$v python bench.py --draws 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); c=d['roofline']['classes_ms_per_step']
print('$VAR=$v', 'evals/s %.4g' % d['value'], ' '.join('%s %.2f' % (k, x) for k, x in c.items()), 'MHz', d['clocks']['sm_mhz'])"
  done
done
