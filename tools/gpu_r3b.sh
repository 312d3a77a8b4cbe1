#!/bin/bash
set +e
OUT=gpurun_out/r3b
mkdir -p $OUT
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -s > $OUT/t_tc.log 2>&1; echo "tc unit rc=$?" | tee $OUT/steps.log
grep -E "passed|failed|Error|error" $OUT/t_tc.log | tail -5
timeout 300 python tools/debug_t128.py 2>&1 | grep -E "use_tc|sample" | head -14
if grep -q "rc=[^0]" $OUT/steps.log; then grep -E "conv_tc " $OUT/t_tc.log | tail -8; exit 0; fi
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "student or faceana_run" 2>&1 | tail -2
for v in 1 0; do SKPS_TCT=$v SKPS_BENCH_OPS=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_tct$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('tct=$v value', d['value'], 'ms', d['ms_per_step'], 'conv2', r['kernel_ms'], r['achieved'], r['frac_of_split_ceiling'])"; grep " op 60 " $OUT/bench_tct$v.err; done
