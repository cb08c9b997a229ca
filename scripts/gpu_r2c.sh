#!/bin/bash
# N-GPU records: correctness + timing of the data-parallel training step, bench lines for c2 / c5
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_step_n$N.log 2>&1; echo "dp rc=$?"; tail -3 gpurun_out/dp_step_n$N.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "bench rc=$?"; cat gpurun_out/bench_c2_n$N.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "bench c5 rc=$?"; cat gpurun_out/bench_c5_n$N.json | cut -c1-400
