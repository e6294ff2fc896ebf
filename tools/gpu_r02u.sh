#!/bin/bash
# parity-sensitive GPU tests + bench after: reciprocal instead of division (blackbody, Poisson terms, power law), statistic prefetch
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02u_pytest.log
tail -n 5 $OUT/r02u_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02u_bench.json 2> $OUT/r02u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['max_rel_value'], d['parity']['max_rel_grad_rownorm'], d['parity']['pass'], d['clocks'])
print(d['roofline']['classes_ms_per_step'])
print({k: v.get('evals_s', v.get('us_per_leapfrog')) for k, v in d['configs'].items()})
print({k: (v['parity']['max_rel_value'], v['parity']['max_rel_grad_rownorm']) for k, v in d['configs'].items()})
PY
tail -n 3 $OUT/r02u_bench.err
