#!/bin/bash
set +e
OUT=gpurun_out/r3c
mkdir -p $OUT
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -s -k "tct" > $OUT/t_tc.log 2>&1; echo "tct unit rc=$?" | tee $OUT/steps.log
grep -E "passed|failed|Error|error|conv_tc \(" $OUT/t_tc.log | tail -8
if grep -q "rc=[^0]" $OUT/steps.log; then exit 0; fi
for v in 1 0; do SKPS_TCT_K3=$v SKPS_BENCH_OPS=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_k3$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('tct_k3=$v value', d['value'], 'ms', d['ms_per_step'], 'conv2', r['kernel_ms'], r['achieved'], r['frac_of_split_ceiling'])"; done
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "student or faceana_run" 2>&1 | tail -2
