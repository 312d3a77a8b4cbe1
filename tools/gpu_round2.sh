#!/bin/bash
# GPU round 2: parity suite, headline bench + A/B switches, launch lists, ncu --set full captures.
set +e
OUT=gpurun_out/$1
mkdir -p $OUT
echo "== new-kernel unit tests (short timeouts; a failing path is switched off for the rest of the run)" | tee $OUT/steps.log
timeout 120 python -m pytest tests/test_conv_tc_gpu.py -q -k "conv_mma" -s > $OUT/t_conv_mma.log 2>&1; MM=$?; echo "conv_mma rc=$MM" | tee -a $OUT/steps.log
grep -E "conv_mma \(|rror" $OUT/t_conv_mma.log | head -8
if [ $MM -ne 0 ]; then export SKPS_CONV_MMA=0; echo "FALLBACK: SKPS_CONV_MMA=0" | tee -a $OUT/steps.log; fi
echo "== gpu parity suite" | tee -a $OUT/steps.log
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_conv_tc_gpu.py::test_conv_mma_small_channel_3x3 > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -4 $OUT/t_gpu.log
grep -E "teacher (cuda|fp32)" $OUT/t_gpu.log
echo "== bench (default)" | tee -a $OUT/steps.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
python -c "
import json; d=json.load(open('$OUT/bench.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['kernel_ms'])"
for sw in ${AB_SWITCHES:-"SKPS_DW_ROWS2=0" "SKPS_SE_FUSE=0"}; do
  env $sw timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$sw.json 2> $OUT/bench_$sw.err
  python -c "
import json; d=json.load(open('$OUT/bench_$sw.json')); print('$sw', 'value', d['value'], 'ms', d['ms_per_step'])" | tee -a $OUT/steps.log
done
echo "== teacher sweep" | tee -a $OUT/steps.log
timeout 900 python tools/bench_teacher.py --batches ${TEACHER_BATCHES:-1,2,4,8,16,32,64,128,256,512,1024} --steps 5 --out $OUT/teacher_sweep.json > $OUT/teacher_sweep.log 2>&1; echo "teacher rc=$?" | tee -a $OUT/steps.log
cut -c 1-160 $OUT/teacher_sweep.log
SKPS_CONV_MMA=0 timeout 600 python tools/bench_teacher.py --batches ${TEACHER_BATCHES:-64,256} --steps 5 > $OUT/teacher_sweep_nomma.log 2>&1; echo "teacher (no conv_mma) rc=$?" | tee -a $OUT/steps.log
cut -c 1-160 $OUT/teacher_sweep_nomma.log
echo "== ncu launch lists" | tee -a $OUT/steps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/teacher_b64_launches.csv python tools/profile_student.py 64 1 teacher > $OUT/ncu_teacher.log 2>&1; echo "ncu teacher rc=$?" | tee -a $OUT/steps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/student_b256_launches.csv python tools/profile_student.py 256 1 student > $OUT/ncu_student.log 2>&1; echo "ncu student rc=$?" | tee -a $OUT/steps.log
if [ -n "$NCU_FULL_DW" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dw_tma|upcat" -c 22 -o $OUT/full_dw python tools/profile_student.py 256 1 student > $OUT/ncu_full_dw.log 2>&1; echo "ncu full dw rc=$?" | tee -a $OUT/steps.log
fi
if [ -n "$NCU_FULL_TC" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 18 -c 3 -o $OUT/full_teacher_tc python tools/profile_student.py 64 1 teacher > $OUT/ncu_full_tc.log 2>&1; echo "ncu full teacher tc rc=$?" | tee -a $OUT/steps.log
fi
echo done | tee -a $OUT/steps.log
