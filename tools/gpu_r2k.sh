#!/bin/bash
set +e
OUT=gpurun_out/r2k
mkdir -p $OUT
bash tools/gpu_cycle.sh r2k
echo "== streams bench" | tee -a $OUT/steps.log
timeout 600 python tools/bench_streams.py --streams 16 --batches 12 > $OUT/streams.jsonl 2> $OUT/streams.err; echo "streams rc=$?" | tee -a $OUT/steps.log
cat $OUT/streams.jsonl; tail -3 $OUT/streams.err
timeout 300 python tools/bench_pipeline.py 40 --no-cpu > $OUT/pipeline_single.jsonl 2> $OUT/pipeline_single.err; echo "single rc=$?" | tee -a $OUT/steps.log
cat $OUT/pipeline_single.jsonl
python tools/launch_table.py $OUT/student_b256_launches.csv > $OUT/launch_table.txt 2>&1; tail -40 $OUT/launch_table.txt
echo done | tee -a $OUT/steps.log
