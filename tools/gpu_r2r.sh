#!/bin/bash
set +e
OUT=gpurun_out/r2r
mkdir -p $OUT
run() { env $1 SKPS_BENCH_OPS=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_$2.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$2 value', d['value'], 'ms', d['ms_per_step'])"; grep " op " $OUT/bench_$2.err | awk '{print $4, $5, $6}' > $OUT/ops_$2.txt; }
run X=1 base
run SKPS_TC_TMA_STORE=0 nostore
run SKPS_TC_MT=1 mt1
run SKPS_FOLD_AFFINE=0 nofold
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "student or faceana_run" 2>&1 | tail -3
paste $OUT/ops_base.txt $OUT/ops_nostore.txt $OUT/ops_mt1.txt | awk '{printf "%3s %-16s base %8s  nostore %8s  mt1 %8s\n", $1, $2, $3, $6, $9}'
