#!/bin/bash
# data-parallel training step record: bash scripts/gpu_dp.sh N [images_per_rank_in_the_check] [model] [prompt] [local_batch]
N=${1:-2}
mkdir -p gpurun_out
tag=${3:-tiny}
DP_IMAGES_PER_RANK=${2:-2} DP_MODEL=$tag DP_PROMPT=${4:-boxes} DP_LOCAL_BATCH=${5:-32} timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_step_${tag}_n$N.log 2>&1; echo "dp rc=$?"; grep -E '^\[\{"rank"|dp check ok|Error' gpurun_out/dp_step_${tag}_n$N.log | cut -c1-1100
