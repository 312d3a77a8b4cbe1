#!/bin/bash
# 4-GPU check of the bench contract (what the driver's scaling run does at N=4)
set +e
OUT=gpurun_out/n8
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench8.json 2> $OUT/bench8.err; echo "bench8 rc=$?" | tee $OUT/steps.log
grep -E "bench [0-9.]+s" $OUT/bench8.err | tail -14
python -c "
import json; d=json.load(open('$OUT/bench8.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value']); print({k:(round(v['frames_per_s']), round(v['faces_per_s']), v['gather_to_rank0']) for k,v in d['pipeline'].items()})"
free -g | head -2; nproc
