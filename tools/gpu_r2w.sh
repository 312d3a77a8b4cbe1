#!/bin/bash
set +e
OUT=gpurun_out/r2w
mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_streams_gpu.py -q -x > $OUT/t_par.log 2>&1; echo "parity rc=$?" | tee $OUT/steps.log
tail -4 $OUT/t_par.log
timeout 300 python tools/bench_detector.py 1 16 > $OUT/det.jsonl 2> $OUT/det.err; cat $OUT/det.jsonl | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/det_b16_launches.csv python tools/profile_student.py 16 1 detector > $OUT/ncu_det.log 2>&1; echo "ncu det rc=$?" | tee -a $OUT/steps.log
python tools/launch_table.py $OUT/det_b16_launches.csv 20 detector > $OUT/det_launch_table.txt 2>&1; head -24 $OUT/det_launch_table.txt; tail -1 $OUT/det_launch_table.txt
timeout 300 python tools/bench_streams.py --streams 16 --batches 12 2>/dev/null | cut -c1-250
