#!/bin/bash
# ncu evidence for the round-2 pipeline: launch list + one full capture per kernel of the bench chunk, and of cfg4
set -u
OUT=gpurun_out
mkdir -p $OUT
SHORT="python bench.py --draws 37888 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 $SHORT > $OUT/r02k_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r02k_launches.csv $SHORT > $OUT/r02k_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unfold|fold_i8|stat_slice' -s 12 -c 5 -f -o $OUT/prof_r02k $SHORT > $OUT/r02k_ncu2.log 2>&1
echo "ncu2 rc=$?"
# cfg4 (band-sparse, DMMA engine): launch list + full capture
timeout 300 python tools/cfg4_ncu_run.py > $OUT/r02k_cfg4_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -s 4 -c 4 -f -o $OUT/prof_r02k_cfg4 python tools/cfg4_ncu_run.py > $OUT/r02k_ncu_cfg4.log 2>&1
echo "ncu cfg4 rc=$?"
# summaries are made on the box: the two reports together exceed what gpurun copies back
ncu -i $OUT/prof_r02k.ncu-rep --page raw --csv > $OUT/r02k_raw.csv 2>/dev/null
ncu -i $OUT/prof_r02k.ncu-rep --page source --csv --print-source sass > $OUT/r02k_src.csv 2>/dev/null
ncu -i $OUT/prof_r02k_cfg4.ncu-rep --page raw --csv > $OUT/r02k_cfg4_raw.csv 2>/dev/null
ncu -i $OUT/prof_r02k_cfg4.ncu-rep --page source --csv --print-source sass > $OUT/r02k_cfg4_src.csv 2>/dev/null
rm -f $OUT/prof_r02k_cfg4.ncu-rep
ls -la $OUT/
