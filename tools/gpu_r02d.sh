#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/r02d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02d_pytest.log
tail -15 $OUT/r02d_pytest.log
