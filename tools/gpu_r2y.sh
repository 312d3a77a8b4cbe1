#!/bin/bash
set +e
OUT=gpurun_out/r2y
mkdir -p $OUT
timeout 300 python -m pytest tests/test_small_kernels_gpu.py -q -x -k "se_fc" 2>&1 | tail -2
for v in 1 0; do SKPS_SE_CLUSTER=$v SKPS_BENCH_OPS=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_se$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('se_cluster=$v value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['op_class_ms']['OP_SE_FC'])"; grep "SE_FC" $OUT/bench_se$v.err | awk '{print $4,$6}' | tr '\n' ' '; echo; done
