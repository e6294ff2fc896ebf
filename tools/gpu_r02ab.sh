#!/bin/bash
# bench with the 1e-5 tolerance side line + the reference arm
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02ab_bench.json 2> $OUT/r02ab_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02ab_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'])
for k, v in d['configs'].items(): print(k, v.get('evals_s'), v.get('us_per_leapfrog'), v['parity'] and {q: v['parity'][q] for q in ('max_rel_value','max_rel_grad_batchnorm','max_rel_grad_rownorm','pass')})
PY
tail -n 3 $OUT/r02ab_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02ab_bench_reference.json 2>> $OUT/r02ab_bench.err; echo "ref rc=$?"; cut -c1-600 $OUT/r02ab_bench_reference.json
