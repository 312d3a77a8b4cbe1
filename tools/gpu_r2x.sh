#!/bin/bash
set +e
OUT=gpurun_out/r2x
mkdir -p $OUT
timeout 300 python -m pytest tests/test_small_kernels_gpu.py -q -x -s -k "se_fc" > $OUT/t_se.log 2>&1; echo "se unit rc=$?" | tee $OUT/steps0.log
tail -4 $OUT/t_se.log
if grep -q "rc=[^0]" $OUT/steps0.log; then grep -E "Error|error|assert" $OUT/t_se.log | head; exit 0; fi
bash tools/gpu_cycle.sh r2x quick
grep "SE_FC" $OUT/bench.err | head -9
