#!/bin/bash
# end-of-round evidence on one GPU: full suite, bench, smoke, then the ncu launch list of the bench chunk and a
# full capture of the small-batch kernel (each after its plain run has exited 0)
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02aq_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02aq_pytest.log
tail -n 8 $OUT/r02aq_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02aq_bench.json 2> $OUT/r02aq_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02aq_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'])
for k, v in d['configs'].items(): print(k, v.get('evals_s'), v.get('us_per_leapfrog'), v.get('us_per_leapfrog_in_callers_cuda_graph'), v['parity'] and v['parity']['pass'])
PY
tail -n 3 $OUT/r02aq_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02aq_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/r02aq_smoke.txt
