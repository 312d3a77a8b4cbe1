#!/bin/bash
set +e
OUT=gpurun_out/r2t
mkdir -p $OUT
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -s > $OUT/t_tc.log 2>&1; echo "tc unit rc=$?" | tee $OUT/steps0.log
grep -E "passed|failed|Error|rel err [0-9.e+-]*$" $OUT/t_tc.log | tail -6
if grep -q "rc=[^0]" $OUT/steps0.log; then exit 0; fi
bash tools/gpu_cycle.sh r2t quick
python tools/launch_table.py $OUT/student_b256_launches.csv > $OUT/launch_table.txt 2>&1; head -3 $OUT/launch_table.txt
