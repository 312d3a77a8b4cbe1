#!/bin/bash
# Round-2 call A: validate the any-width tcgen05 tiles on the detector, capture the dominant conv with ncu --set full,
# today's baseline bench.
set +e
OUT=gpurun_out/r2a
mkdir -p $OUT
echo "== detector tests, SKPS_TC_ANY_W=1" | tee $OUT/steps.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/t_det_anyw.log 2>&1; echo "rc=$?" | tee -a $OUT/steps.log
tail -5 $OUT/t_det_anyw.log
echo "== detector timing" | tee -a $OUT/steps.log
SKPS_TC_ANY_W=0 timeout 300 python tools/bench_detector.py 1 16 > $OUT/det_anyw0.jsonl 2> $OUT/det_anyw0.err; echo "rc=$?" | tee -a $OUT/steps.log
SKPS_TC_ANY_W=1 timeout 300 python tools/bench_detector.py 1 16 > $OUT/det_anyw1.jsonl 2> $OUT/det_anyw1.err; echo "rc=$?" | tee -a $OUT/steps.log
cat $OUT/det_anyw0.jsonl $OUT/det_anyw1.jsonl
echo "== bench" | tee -a $OUT/steps.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/steps.log
cut -c 1-400 $OUT/bench.json
echo "== ncu full, dominant conv" | tee -a $OUT/steps.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_conv2 python tools/profile_op.py dominant 256 1 > $OUT/ncu_full_conv2.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
tail -2 $OUT/ncu_full_conv2.log
echo done | tee -a $OUT/steps.log
