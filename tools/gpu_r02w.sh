#!/bin/bash
# full suite + bench + small-batch latency after the one-kernel small-batch path
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02w_pytest.log
tail -n 12 $OUT/r02w_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02w_bench.json 2> $OUT/r02w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02w_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'])
print(d['roofline']['classes_ms_per_step'])
print({k: v.get('evals_s', v.get('us_per_leapfrog')) for k, v in d['configs'].items()})
print(d['configs']['cfg3']['us_per_leapfrog'], d['configs']['cfg3']['engine'])
PY
tail -n 3 $OUT/r02w_bench.err
python tools/small_batch_latency.py > $OUT/r02w_small_batch.txt 2>&1; tail -n 12 $OUT/r02w_small_batch.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02w_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/r02w_smoke.txt
