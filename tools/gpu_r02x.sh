#!/bin/bash
# ncu evidence for the kernels of the second half of round 2: launch list + one full capture per kernel of the
# bench chunk, and of the one-kernel small-batch path (cfg 3)
set -u
OUT=gpurun_out
mkdir -p $OUT
SHORT="python bench.py --draws 37888 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 $SHORT > $OUT/r02x_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r02x_launches.csv $SHORT > $OUT/r02x_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unfold|fold_i8|stat_slice' -s 12 -c 5 -f -o $OUT/prof_r02x $SHORT > $OUT/r02x_ncu2.log 2>&1
echo "ncu2 rc=$?"
timeout 300 python tools/tiny_ncu_run.py > $OUT/r02x_tiny_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tiny_eval' -s 8 -c 4 -f -o $OUT/prof_r02x_tiny python tools/tiny_ncu_run.py > $OUT/r02x_ncu_tiny.log 2>&1
echo "ncu tiny rc=$?"
ncu -i $OUT/prof_r02x.ncu-rep --page raw --csv > $OUT/r02x_raw.csv 2>/dev/null
ncu -i $OUT/prof_r02x.ncu-rep --page source --csv --print-source sass > $OUT/r02x_src.csv 2>/dev/null
ncu -i $OUT/prof_r02x_tiny.ncu-rep --page raw --csv > $OUT/r02x_tiny_raw.csv 2>/dev/null
rm -f $OUT/prof_r02x_tiny.ncu-rep
ls -la $OUT/ | tail -12
