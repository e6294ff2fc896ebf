#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fold_engines.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $OUT/r02g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02g_pytest.log
tail -5 $OUT/r02g_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs > $OUT/r02g_bench.json 2> $OUT/r02g_bench.err; echo "bench rc=$?"
ELISA_B200_I8_SKIP=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-e2e > $OUT/r02g_bench_noskip.json 2> $OUT/r02g_bench_noskip.err; echo "noskip rc=$?"
python - <<'PY'
import json
for f in ('r02g_bench', 'r02g_bench_noskip'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, d['value'], d['roofline']['classes_ms_per_step'], d['roofline'].get('digit_products'), d['roofline']['frac'], d.get('parity', {}) and d['parity'].get('pass'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -3 $OUT/r02g_bench.err
