#!/bin/bash
# Round-end validation on one B200: full GPU parity suite, smoke(), the headline bench exactly as the driver runs it,
# the reference arm, the Teacher batch sweep (config 4), the full-pipeline configs 3/5, launch lists and ncu captures.
set +e
OUT=gpurun_out/$1
mkdir -p $OUT
echo "== gpu parity suite" | tee $OUT/steps.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -3 $OUT/t_gpu.log
grep -E "teacher (cuda|fp32)" $OUT/t_gpu.log
echo "== smoke" | tee -a $OUT/steps.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/steps.log; tail -1 $OUT/smoke.log
echo "== bench (driver defaults)" | tee -a $OUT/steps.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
python -c "
import json; d=json.load(open('$OUT/bench.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['kernel_ms'], 'cpu', d['cpu_baseline']['value'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?" | tee -a $OUT/steps.log
cut -c 1-200 $OUT/bench_reference.json
echo "== teacher sweep" | tee -a $OUT/steps.log
timeout 900 python tools/bench_teacher.py --steps 5 --out $OUT/teacher_sweep.json > $OUT/teacher_sweep.log 2>&1; echo "teacher rc=$?" | tee -a $OUT/steps.log
cut -c 1-150 $OUT/teacher_sweep.log
echo "== pipeline configs 3 and 5 (one GPU)" | tee -a $OUT/steps.log
timeout 600 python tools/bench_pipeline.py 40 > $OUT/pipeline.log 2> $OUT/pipeline.err; echo "pipeline rc=$?" | tee -a $OUT/steps.log
cut -c 1-260 $OUT/pipeline.log
echo "== ncu" | tee -a $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/student_b256_launches.csv python tools/profile_student.py 256 1 student > $OUT/ncu_student.log 2>&1; echo "ncu student rc=$?" | tee -a $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/teacher_b64_launches.csv python tools/profile_student.py 64 1 teacher > $OUT/ncu_teacher.log 2>&1; echo "ncu teacher rc=$?" | tee -a $OUT/steps.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_mma_kernel" -s 2 -c 2 -o $OUT/full_conv_mma python tools/profile_student.py 64 1 teacher > $OUT/ncu_full_mma.log 2>&1; echo "ncu full conv_mma rc=$?" | tee -a $OUT/steps.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 36 -c 2 -o $OUT/full_conv2 python tools/profile_student.py 256 1 student > $OUT/ncu_full_conv2.log 2>&1; echo "ncu full conv2/hm rc=$?" | tee -a $OUT/steps.log
echo done | tee -a $OUT/steps.log
