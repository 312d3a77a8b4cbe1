#!/bin/bash
# Round-2 call C: why is the fused depthwise->1x1 kernel slow?  ncu --set full with source-level stalls.
set +e
OUT=gpurun_out/r2c
mkdir -p $OUT
echo "== gpu suite" | tee $OUT/steps.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -6 $OUT/t_gpu.log
echo "== ncu full: fused up2 head" | tee -a $OUT/steps.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_up2 python tools/profile_op.py upsampler2/conv1 256 1 > $OUT/ncu_up2.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
tail -2 $OUT/ncu_up2.log
echo "== ncu full: fused blocks.0.0" | tee -a $OUT/steps.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_b00 python tools/profile_op.py blocks.0.0/conv_dw 256 1 > $OUT/ncu_b00.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
tail -2 $OUT/ncu_b00.log
echo "== detector" | tee -a $OUT/steps.log
timeout 300 python tools/bench_detector.py 1 16 > $OUT/det.jsonl 2> $OUT/det.err; echo "det rc=$?" | tee -a $OUT/steps.log
cat $OUT/det.jsonl
SKPS_XF=0 timeout 300 python tools/bench_detector.py 1 16 > $OUT/det_noxf.jsonl 2> $OUT/det_noxf.err; echo "det rc=$?" | tee -a $OUT/steps.log
cat $OUT/det_noxf.jsonl
echo done | tee -a $OUT/steps.log
