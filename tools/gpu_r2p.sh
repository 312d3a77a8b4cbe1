#!/bin/bash
# 2-GPU check of the bench contract (torchrun, pipeline leg with the NCCL gather)
set +e
OUT=gpurun_out/r2p
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench2.json 2> $OUT/bench2.err; echo "bench2 rc=$?" | tee $OUT/steps.log
tail -5 $OUT/bench2.err
python -c "
import json; d=json.load(open('$OUT/bench2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print({k:(round(v['frames_per_s']), round(v['faces_per_s']), v['gather_to_rank0']) for k,v in d['pipeline'].items()})"
timeout 600 python -m pytest tests/test_teacher.py -q -x -s -k "128_cuda or 192" 2>&1 | grep -E "teacher@128|passed|failed"
