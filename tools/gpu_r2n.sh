#!/bin/bash
set +e
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -s > $OUT/t_tc.log 2>&1; echo "tc unit rc=$?" | tee $OUT/steps0.log
grep -E "conv_tc|passed|failed|Error" $OUT/t_tc.log | tail -40
if grep -q "rc=[^0]" $OUT/steps0.log; then
  SKPS_TC_K3=0 timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -s -k "halo" 2>&1 | tail -5
  exit 0
fi
bash tools/gpu_cycle.sh r2n quick
python tools/launch_table.py $OUT/student_b256_launches.csv > $OUT/launch_table.txt 2>&1; head -12 $OUT/launch_table.txt
