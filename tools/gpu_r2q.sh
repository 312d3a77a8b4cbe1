#!/bin/bash
set +e
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_small python tools/profile_op.py "#2,16,37,40" 256 1 > $OUT/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
python tools/ncu_summary.py $OUT/full_small.ncu-rep > $OUT/summary.txt 2>&1
for k in 1 2 3 4; do python tools/ncu_stalls.py $OUT/full_small.ncu-rep ":::$k" 30 > $OUT/stalls_$k.txt 2>&1; done
echo done | tee -a $OUT/steps.log
