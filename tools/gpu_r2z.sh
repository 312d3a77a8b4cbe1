#!/bin/bash
set +e
OUT=gpurun_out/r2z
mkdir -p $OUT
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_xf_gpu.py -q -x > $OUT/t_unit.log 2>&1; echo "unit rc=$?" | tee $OUT/steps.log
tail -2 $OUT/t_unit.log
if grep -q "rc=[^0]" $OUT/steps.log; then exit 0; fi
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -3 $OUT/t_gpu.log
for v in 1 0; do SKPS_PDL=$v timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline 2> $OUT/bench_pdl$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl=$v value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'det', [(x['batch'], round(x['ms'],3)) for x in (d.get('detector') or [])])"; done
for v in 1 0; do echo "pdl=$v"; SKPS_PDL=$v timeout 300 python tools/bench_detector.py 1 16 2>/dev/null | cut -c1-140; SKPS_PDL=$v timeout 300 python tools/bench_streams.py --streams 16 --batches 12 2>/dev/null | cut -c1-150; SKPS_PDL=$v timeout 200 python tools/bench_pipeline.py 40 --no-cpu 2>/dev/null | cut -c1-150; done
