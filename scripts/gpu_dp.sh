#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
DP_IMAGES_PER_RANK=${2:-2} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_step_n$N.log 2>&1; echo "dp rc=$?"; grep -E '^\[\{"rank"|dp check ok|Error' gpurun_out/dp_step_n$N.log | cut -c1-900
