#!/bin/bash
set +e
OUT=gpurun_out/r2f
mkdir -p $OUT
bash tools/gpu_cycle.sh r2f quick
echo "== ncu full: fused up2 head" | tee -a $OUT/steps.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_up2 python tools/profile_op.py upsampler2/conv1 256 1 > $OUT/ncu_up2.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
echo "== streams bench" | tee -a $OUT/steps.log
timeout 600 python tools/bench_streams.py --streams 16 --batches 12 > $OUT/streams.jsonl 2> $OUT/streams.err; echo "streams rc=$?" | tee -a $OUT/steps.log
cat $OUT/streams.jsonl; tail -3 $OUT/streams.err
timeout 300 python tools/bench_pipeline.py 40 --no-cpu > $OUT/pipeline_single.jsonl 2> $OUT/pipeline_single.err; echo "single rc=$?" | tee -a $OUT/steps.log
cat $OUT/pipeline_single.jsonl
echo "== teacher path experiments" | tee -a $OUT/steps.log
for v in "X=1" "SKPS_CONV_MMA=0" "SKPS_XF=0" "SKPS_TC_SMALL=0" "SKPS_TC_STRIDE2=0" "SKPS_TC_ANY_W=0"; do
  env $v timeout 300 python -m pytest tests/test_teacher.py -q -s -k "fp32_fallback_path" 2>&1 | grep -E "teacher fp32|passed|failed" | sed "s/^/[$v] /"
done | tee $OUT/teacher_variants.log
echo done | tee -a $OUT/steps.log
