#!/bin/bash
# default chunk = two row tiles per SM: the tests that run full shapes / both engines / the guard / the re-fit, then the bench
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 230 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fold_engines.py tests/test_gpu_guard.py tests/test_gpu_fit.py tests/test_gpu_jacobian.py -m gpu -q -x 2>&1 | tail -4
timeout 90 python bench.py --steps 5 --warmup 3 > $OUT/r02av_bench.json 2> $OUT/r02av_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02av_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'], d['config']['chunk'], d['roofline']['frac'], d['roofline']['traffic'])
print(d['roofline']['classes_ms_per_step'])
for k, v in d['configs'].items(): print(k, v.get('evals_s'), v.get('us_per_leapfrog'), v['parity'] and v['parity']['pass'])
PY
tail -n 2 $OUT/r02av_bench.err
