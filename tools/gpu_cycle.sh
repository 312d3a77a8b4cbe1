#!/bin/bash
# Generic round-2 GPU cycle: bash tools/gpu_cycle.sh <tag> [quick]
#   new-kernel unit tests (separate processes, short timeouts) -> full GPU suite -> bench -> launch list
set +e
OUT=gpurun_out/$1
mkdir -p $OUT
echo "== unit tests" | tee $OUT/steps.log
timeout 300 python -m pytest tests/test_conv_xf_gpu.py -q -x -s > $OUT/t_unit.log 2>&1; echo "unit rc=$?" | tee -a $OUT/steps.log
tail -4 $OUT/t_unit.log
if grep -q "unit rc=[^0]" $OUT/steps.log; then grep -E "^E  |rel err" $OUT/t_unit.log | head -20; exit 0; fi
echo "== gpu suite" | tee -a $OUT/steps.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -6 $OUT/t_gpu.log
echo "== bench" | tee -a $OUT/steps.log
SKPS_BENCH_OPS=1 timeout 600 python bench.py --steps 20 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
python -c "
import json; d=json.load(open('$OUT/bench.json')); r=d['roofline']; print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']); print('dominant', r['kernel_ms'], r['achieved'], 'hbm', r['hbm_kernel']); print(r['op_class_ms'], r['op_sum_ms']); print('pipeline', {k: (round(v['frames_per_s']), round(v['faces_per_s'])) for k, v in (d.get('pipeline') or {}).items()}); print('detector', [(x['batch'], round(x['ms'], 3)) for x in (d.get('detector') or [])])"
grep " op " $OUT/bench.err | sort -k6 -n -r | head -30
tail -3 $OUT/bench.err
echo "== launch list" | tee -a $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/student_b256_launches.csv python tools/profile_student.py 256 1 student > $OUT/ncu_student.log 2>&1; echo "ncu student rc=$?" | tee -a $OUT/steps.log
if [ "$2" != "quick" ]; then
echo "== detector" | tee -a $OUT/steps.log
timeout 300 python tools/bench_detector.py 1 16 > $OUT/det.jsonl 2> $OUT/det.err; echo "det rc=$?" | tee -a $OUT/steps.log
cat $OUT/det.jsonl
fi
echo done | tee -a $OUT/steps.log
