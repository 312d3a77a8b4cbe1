#!/bin/bash
# One gpurun call: kernel unit tests first (under timeouts, with fallbacks switched on if a new path fails),
# then the GPU parity suite, the headline bench, the Teacher sweep and the ncu launch lists.
# Everything lands in gpurun_out/.
set +e
OUT=gpurun_out/$1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
echo "== conv_tc unit tests" | tee $OUT/steps.log
timeout 420 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "not stride2 and not small_maps" > $OUT/t_convtc_base.log 2>&1; echo "base rc=$?" | tee -a $OUT/steps.log
timeout 240 python -m pytest tests/test_conv_tc_gpu.py -q -k "stride2" -s > $OUT/t_convtc_stride2.log 2>&1; S2=$?; echo "stride2 rc=$S2" | tee -a $OUT/steps.log
timeout 240 python -m pytest tests/test_conv_tc_gpu.py -q -k "small_maps" -s > $OUT/t_convtc_small.log 2>&1; SM=$?; echo "small rc=$SM" | tee -a $OUT/steps.log
if [ $S2 -ne 0 ]; then export SKPS_TC_STRIDE2=0; echo "FALLBACK: SKPS_TC_STRIDE2=0" | tee -a $OUT/steps.log; fi
if [ $SM -ne 0 ]; then export SKPS_TC_SMALL=0; echo "FALLBACK: SKPS_TC_SMALL=0" | tee -a $OUT/steps.log; fi
echo "== gpu parity suite" | tee -a $OUT/steps.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_conv_tc_gpu.py > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -5 $OUT/t_gpu.log
echo "== bench" | tee -a $OUT/steps.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
cat $OUT/bench.json | head -c 600
echo "== teacher sweep" | tee -a $OUT/steps.log
timeout 900 python tools/bench_teacher.py --batches ${TEACHER_BATCHES:-1,8,64,256} --steps 5 --out $OUT/teacher_sweep.json > $OUT/teacher_sweep.log 2>&1; echo "teacher rc=$?" | tee -a $OUT/steps.log
cat $OUT/teacher_sweep.log | cut -c 1-260
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch lists" | tee -a $OUT/steps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/teacher_b64_launches.csv python tools/profile_student.py 64 1 teacher > $OUT/ncu_teacher.log 2>&1; echo "ncu teacher rc=$?" | tee -a $OUT/steps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/student_b256_launches.csv python tools/profile_student.py 256 1 student > $OUT/ncu_student.log 2>&1; echo "ncu student rc=$?" | tee -a $OUT/steps.log
fi
if [ -n "$NCU_FULL" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU_FULL" -c ${NCU_FULL_COUNT:-3} -o $OUT/full_$2 python tools/profile_student.py 256 1 student > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/steps.log
fi
echo done | tee -a $OUT/steps.log
