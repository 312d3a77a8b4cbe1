#!/bin/bash
# Round-2 end validation on one B200: bash tools/gpu_final_r2.sh <tag>
# full GPU parity suite, smoke(), the headline bench exactly as the driver runs it (incl. pipeline + detector legs and the CPU
# baseline), launch lists (student b256, detector b16, teacher b64), ncu --set full of the dominant conv (+ traffic json),
# teacher sweep.
set +e
OUT=gpurun_out/$1
mkdir -p $OUT
echo "== gpu parity suite" | tee $OUT/steps.log
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -3 $OUT/t_gpu.log
grep -E "teacher (cuda|fp32)" $OUT/t_gpu.log
echo "== smoke" | tee -a $OUT/steps.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/steps.log; tail -1 $OUT/smoke.log
echo "== bench (driver defaults)" | tee -a $OUT/steps.log
SKPS_BENCH_OPS=1 timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
python -c "
import json; d=json.load(open('$OUT/bench.json')); r=d['roofline']; print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', r['achieved'], r['kernel_ms'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores']); print('pipeline', {k: (round(v['frames_per_s']), round(v['faces_per_s'])) for k, v in d['pipeline'].items()}); print('detector', [(x['batch'], round(x['ms'], 3)) for x in d['detector']]); print(r['op_class_ms']); print('clocks', d['clocks'])"
echo "== launch lists" | tee -a $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/student_b256_launches.csv python tools/profile_student.py 256 1 student > $OUT/ncu_student.log 2>&1; echo "ncu student rc=$?" | tee -a $OUT/steps.log
python tools/launch_table.py $OUT/student_b256_launches.csv 70 > $OUT/student_launch_table.txt 2>&1; head -8 $OUT/student_launch_table.txt; tail -1 $OUT/student_launch_table.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/det_b16_launches.csv python tools/profile_student.py 16 1 detector > $OUT/ncu_det.log 2>&1; echo "ncu det rc=$?" | tee -a $OUT/steps.log
python tools/launch_table.py $OUT/det_b16_launches.csv 95 detector > $OUT/det_launch_table.txt 2>&1; tail -1 $OUT/det_launch_table.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/teacher_b64_launches.csv python tools/profile_student.py 64 1 teacher > $OUT/ncu_teacher.log 2>&1; echo "ncu teacher rc=$?" | tee -a $OUT/steps.log
python tools/launch_table.py $OUT/teacher_b64_launches.csv 30 teacher > $OUT/teacher_launch_table.txt 2>&1; tail -1 $OUT/teacher_launch_table.txt
echo "== ncu full: dominant conv, stem block, heat-map head" | tee -a $OUT/steps.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_final python tools/profile_op.py "upsampler2/conv2/conv2.0/Conv" 256 1; timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_stem_hm python tools/profile_op.py "#0" 256 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/steps.log
python tools/ncu_summary.py $OUT/full_final.ncu-rep > $OUT/full_final_summary.txt 2>&1
grep -E "Kernel Name|time_duration|dram__bytes|lts__throughput|tensor_cycles" $OUT/full_final_summary.txt | cut -c1-130
echo "== teacher sweep" | tee -a $OUT/steps.log
timeout 900 python tools/bench_teacher.py --steps 5 --out $OUT/teacher_sweep.json > $OUT/teacher_sweep.log 2>&1; echo "teacher rc=$?" | tee -a $OUT/steps.log
cut -c 1-120 $OUT/teacher_sweep.log | tail -11
echo done | tee -a $OUT/steps.log
