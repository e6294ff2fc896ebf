#!/bin/bash
# edge-wise gradient contraction: parity tests of the gradient paths, then A/B against the dual-number contraction
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_fold_engines.py tests/test_gpu_fullsize.py tests/test_gpu_autograd.py -m gpu -x -q > $OUT/r02r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02r_pytest.log
tail -n 5 $OUT/r02r_pytest.log
bash tools/ab_env.sh ELISA_B200_CONTRACT generic edges 2 > $OUT/r02r_ab_contract.txt 2>&1
cat $OUT/r02r_ab_contract.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02r_bench.json 2> $OUT/r02r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['max_rel_grad_rownorm'], d['parity']['pass'])
print(d['roofline']['classes_ms_per_step'])
print({k: v.get('evals_s', v.get('us_per_leapfrog')) for k, v in d['configs'].items()})
print({k: v['parity']['max_rel_grad_rownorm'] for k, v in d['configs'].items()})
PY
tail -n 3 $OUT/r02r_bench.err
