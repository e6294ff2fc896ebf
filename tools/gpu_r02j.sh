#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -x > $OUT/r02j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02j_pytest.log
tail -6 $OUT/r02j_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02j_bench.json 2> $OUT/r02j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print(d['value'], d['e2e'], d['roofline']['classes_ms_per_step'], d['roofline']['frac'], d['parity']['pass'])
print({k: (v.get('evals_s'), v.get('us_per_leapfrog'), v['parity'] and v['parity']['pass']) for k, v in d['configs'].items()})
PY
tail -3 $OUT/r02j_bench.err
