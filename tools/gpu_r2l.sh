#!/bin/bash
# ncu --set full of the five largest non-conv2 launches of the student step (one process, one capture per op)
set +e
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_ops python tools/profile_op.py "#0,59,38,39,61" 256 1 > $OUT/ncu_ops.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
tail -6 $OUT/ncu_ops.log
python tools/ncu_summary.py $OUT/full_ops.ncu-rep > $OUT/summary.txt 2>&1
for k in 0 1 2 3 4; do python tools/ncu_stalls.py $OUT/full_ops.ncu-rep ":::$k" 25 > $OUT/stalls_$k.txt 2>&1; done
ls -la $OUT
echo done | tee -a $OUT/steps.log
