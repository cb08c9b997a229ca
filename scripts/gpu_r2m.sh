#!/bin/bash
# ncu launch list of the final code (per-launch time and DRAM bytes), after a plain run of the same command
mkdir -p gpurun_out
timeout 60 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 90 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ph_|seg_sort|match_|loss_kernel|grad_kernel|scale_kernel|unpack' -c 30 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_c2.csv
