#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_fold_engines.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_conv.py tests/test_gpu_random.py -m gpu -q -x > $OUT/r02i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02i_pytest.log
tail -5 $OUT/r02i_pytest.log
bash tools/ab_env.sh ELISA_B200_GRAD contract planes 3 2>&1 | tee $OUT/r02i_ab_grad.txt
