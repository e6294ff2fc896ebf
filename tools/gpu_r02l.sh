#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -x > $OUT/r02l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02l_pytest.log
tail -6 $OUT/r02l_pytest.log
timeout 600 python tools/csr_race.py > $OUT/r02l_csr_race.md 2> $OUT/r02l_csr_race.err; echo "race rc=$?"; cat $OUT/r02l_csr_race.md; tail -3 $OUT/r02l_csr_race.err
timeout 600 python tools/config_rates.py > $OUT/r02l_config_rates.txt 2>&1; cat $OUT/r02l_config_rates.txt
timeout 300 python tools/fit_rate.py > $OUT/r02l_fit_rate.txt 2>&1; tail -3 $OUT/r02l_fit_rate.txt
