#!/bin/bash
set +e
OUT=gpurun_out/r3e
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -s > $OUT/t.log 2>&1; echo "tests rc=$?" | tee $OUT/steps.log
grep -E "teacher@128|passed|failed|Error" $OUT/t.log | tail -5
timeout 300 python tools/bench_streams.py --streams 16 --batches 12 2>/dev/null | cut -c1-200
for v in 1 0; do SKPS_GAP_SSE=$v SKPS_BENCH_OPS=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_g$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gap_sse=$v value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"; grep -E "GAP|SE_FC|SCSE" $OUT/bench_g$v.err | tail -4 | awk '{print $4,$5,$6}' | tr '\n' ' '; echo; done
