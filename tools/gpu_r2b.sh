#!/bin/bash
# Round-2 call B: fused producer->1x1 conv kernels (conv_xf.cu): unit tests in separate processes, then the suite + bench.
set +e
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== xf unit tests" | tee $OUT/steps.log
for k in scale_matches depthwise_pointwise upsample_concat magnitudes; do
  timeout 240 python -m pytest tests/test_conv_xf_gpu.py -q -s -k $k > $OUT/t_xf_$k.log 2>&1; echo "xf $k rc=$?" | tee -a $OUT/steps.log
  grep -E "rel err|passed|failed|Error|error" $OUT/t_xf_$k.log | tail -12
done
if grep -q "rc=[^0]" $OUT/steps.log; then echo "unit tests failed: skipping the rest" | tee -a $OUT/steps.log; exit 0; fi
echo "== gpu suite" | tee -a $OUT/steps.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/t_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $OUT/steps.log
tail -8 $OUT/t_gpu.log
echo "== bench" | tee -a $OUT/steps.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
cut -c 1-300 $OUT/bench.json
echo "== launch list" | tee -a $OUT/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/student_b256_launches.csv python tools/profile_student.py 256 1 student > $OUT/ncu_student.log 2>&1; echo "ncu student rc=$?" | tee -a $OUT/steps.log
echo "== detector" | tee -a $OUT/steps.log
timeout 300 python tools/bench_detector.py 1 16 > $OUT/det.jsonl 2> $OUT/det.err; echo "det rc=$?" | tee -a $OUT/steps.log
cat $OUT/det.jsonl
echo done | tee -a $OUT/steps.log
