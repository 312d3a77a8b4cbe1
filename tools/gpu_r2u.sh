#!/bin/bash
set +e
OUT=gpurun_out/r2u
mkdir -p $OUT
echo "== detector launch list (batch 16)" | tee $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/det_b16_launches.csv python tools/profile_student.py 16 1 detector > $OUT/ncu_det.log 2>&1; echo "ncu det rc=$?" | tee -a $OUT/steps.log
python tools/launch_table.py $OUT/det_b16_launches.csv 40 detector > $OUT/det_launch_table.txt 2>&1; head -45 $OUT/det_launch_table.txt
echo "== teacher variants" | tee -a $OUT/steps.log
for v in "X=1" "SKPS_CONV_MMA=0"; do
  env $v timeout 300 python -m pytest tests/test_teacher.py -q -s -k "cuda" 2>&1 | grep -E "teacher (cuda|fp32)|passed|failed" | sed "s/^/[$v] /"
done | tee $OUT/teacher_variants.log
echo "== teacher sweep" | tee -a $OUT/steps.log
timeout 600 python tools/bench_teacher.py --batches 1,16,64,256,1024 --out $OUT/teacher_sweep.json > $OUT/teacher.log 2>&1; tail -6 $OUT/teacher.log | cut -c1-250
echo done | tee -a $OUT/steps.log
