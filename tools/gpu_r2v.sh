#!/bin/bash
set +e
OUT=gpurun_out/r2v
mkdir -p $OUT
bash tools/gpu_cycle.sh r2v
python tools/launch_table.py $OUT/student_b256_launches.csv > $OUT/launch_table.txt 2>&1; head -4 $OUT/launch_table.txt
echo "== detector launch list (batch 16)" | tee -a $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/det_b16_launches.csv python tools/profile_student.py 16 1 detector > $OUT/ncu_det.log 2>&1; echo "ncu det rc=$?" | tee -a $OUT/steps.log
python tools/launch_table.py $OUT/det_b16_launches.csv 14 detector > $OUT/det_launch_table.txt 2>&1; head -18 $OUT/det_launch_table.txt; tail -1 $OUT/det_launch_table.txt
