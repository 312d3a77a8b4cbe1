#!/bin/bash
set +e
OUT=gpurun_out/r2m
mkdir -p $OUT
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -s -k "conv_hm" > $OUT/t_hm.log 2>&1; echo "hm unit rc=$?" | tee $OUT/steps0.log
tail -5 $OUT/t_hm.log
bash tools/gpu_cycle.sh r2m quick
python tools/launch_table.py $OUT/student_b256_launches.csv > $OUT/launch_table.txt 2>&1; tail -34 $OUT/launch_table.txt
