#!/bin/bash
set +e
OUT=gpurun_out/r2j
mkdir -p $OUT
bash tools/gpu_cycle.sh r2j quick
echo "== ncu full: stem block" | tee -a $OUT/steps.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_stem python tools/profile_op.py stem_block 256 1 > $OUT/ncu_stem.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/steps.log
echo done | tee -a $OUT/steps.log
