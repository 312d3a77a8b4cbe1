#!/bin/bash
# HEAD sanity on one B200: full GPU suite, smoke(), default bench
set +e
OUT=gpurun_out/head
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee $OUT/steps.log; tail -2 $OUT/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
python -c "
import json; d=json.load(open('$OUT/bench.json')); r=d['roofline']; print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', r['achieved'], r['frac'], r['traffic'], 'cpu', d['cpu_baseline']['value']); print('pipeline', {k: (round(v['frames_per_s']), round(v['faces_per_s'])) for k, v in d['pipeline'].items()}); print('launches', d['gpu_launches'], 'clocks', d['clocks'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
