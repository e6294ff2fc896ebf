#!/bin/bash
# end-of-round evidence on one GPU: full suite, bench, smoke, then the ncu launch list of the bench chunk and a
# full capture of the small-batch kernel (each after its plain run has exited 0)
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r02am_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02am_pytest.log
tail -n 8 $OUT/r02am_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02am_bench.json 2> $OUT/r02am_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02am_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['pass'])
for k, v in d['configs'].items(): print(k, v.get('evals_s'), v.get('us_per_leapfrog'), v.get('us_per_leapfrog_in_callers_cuda_graph'), v['parity'] and v['parity']['pass'])
PY
tail -n 3 $OUT/r02am_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02am_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/r02am_smoke.txt
SHORT="python bench.py --draws 37888 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 $SHORT > $OUT/r02am_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r02am_launches.csv $SHORT > $OUT/r02am_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 300 python tools/tiny_ncu_run.py > $OUT/r02am_tiny_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tiny_eval' -s 2 -c 2 -f -o $OUT/prof_r02am_tiny python tools/tiny_ncu_run.py > $OUT/r02am_ncu_tiny.log 2>&1
echo "ncu tiny rc=$?"
ncu -i $OUT/prof_r02am_tiny.ncu-rep --page raw --csv > $OUT/r02am_tiny_raw.csv 2>/dev/null
rm -f $OUT/prof_r02am_tiny.ncu-rep
ls -la $OUT/ | grep r02am
