#!/bin/bash
# one-kernel small-batch path + edge-contraction tests, then the full suite and the bench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_edges.py -m gpu -q > $OUT/r02v_new.log 2>&1; echo "new tests rc=$?" | tee -a $OUT/r02v_new.log
tail -n 30 $OUT/r02v_new.log
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02v_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02v_pytest.log
tail -n 5 $OUT/r02v_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02v_bench.json 2> $OUT/r02v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'])
print(d['roofline']['classes_ms_per_step'])
print({k: v.get('evals_s', v.get('us_per_leapfrog')) for k, v in d['configs'].items()})
print(d['configs']['cfg3'])
PY
tail -n 3 $OUT/r02v_bench.err
python tools/small_batch_latency.py > $OUT/r02v_small_batch.txt 2>&1; tail -n 12 $OUT/r02v_small_batch.txt
