#!/bin/bash
# Two-GPU check (one box): config 5 pipeline sharded over ranks with the NCCL result gather, then the headline bench.
set +e
OUT=gpurun_out/$1
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/smi.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_pipeline.py 30 --configs 4k_16faces --gather > $OUT/pipeline_2gpu.log 2> $OUT/pipeline_2gpu.err; echo "pipeline 2gpu rc=$?" | tee $OUT/steps.log
cut -c 1-300 $OUT/pipeline_2gpu.log; tail -3 $OUT/pipeline_2gpu.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench 2gpu rc=$?" | tee -a $OUT/steps.log
cut -c 1-260 $OUT/bench_2gpu.json
echo done | tee -a $OUT/steps.log
