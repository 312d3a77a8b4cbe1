#!/bin/bash
set +e
OUT=gpurun_out/r2o
mkdir -p $OUT
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_conv2_k3 python tools/profile_op.py dominant 256 1 > $OUT/ncu_k3.log 2>&1; echo "ncu k3 rc=$?" | tee -a $OUT/steps.log
SKPS_TC_K3=0 timeout 600 ncu --profile-from-start off --set full --clock-control none -o $OUT/full_conv2_tap python tools/profile_op.py dominant 256 1 > $OUT/ncu_tap.log 2>&1; echo "ncu tap rc=$?" | tee -a $OUT/steps.log
python tools/ncu_summary.py $OUT/full_conv2_k3.ncu-rep > $OUT/summary_k3.txt 2>&1
python tools/ncu_summary.py $OUT/full_conv2_tap.ncu-rep > $OUT/summary_tap.txt 2>&1
# stem TC variant: parity + time
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "student" > $OUT/t_student.log 2>&1; echo "student parity rc=$?" | tee -a $OUT/steps.log
tail -5 $OUT/t_student.log
for v in 1 0; do SKPS_STEM_TC=$v SKPS_BENCH_OPS=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_stem$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stem_tc=$v value', d['value'], 'ms', d['ms_per_step'])"; grep " op  0 " $OUT/bench_stem$v.err; done
echo done | tee -a $OUT/steps.log
