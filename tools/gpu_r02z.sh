#!/bin/bash
# free energy shifts + the merged one-launch small-batch path: new tests first, then the full suite, bench, smoke
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_free_shift.py tests/test_gpu_tiny.py -m gpu -q -x > $OUT/r02z_new.log 2>&1; echo "new rc=$?" | tee -a $OUT/r02z_new.log
tail -n 40 $OUT/r02z_new.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02z_pytest.log
tail -n 12 $OUT/r02z_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02z_bench.json 2> $OUT/r02z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'])
print(d['roofline']['classes_ms_per_step'])
print({k: v.get('evals_s', v.get('us_per_leapfrog')) for k, v in d['configs'].items()})
PY
tail -n 3 $OUT/r02z_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02z_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/r02z_smoke.txt
