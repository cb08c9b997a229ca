#!/bin/bash
# ncu evidence of the current code: (1) launch list of the bench command, (2) full capture of the persistence kernel on 148 pred maps
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ph_|seg_sort|match_|loss_kernel|grad_kernel' -c 60 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_c2.csv
timeout 120 python scripts/ncu_ph.py pred > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ph_small -c 1 -o gpurun_out/ph_small_r2 -f python scripts/ncu_ph.py pred > gpurun_out/ncu_run.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/ph_small_r2.ncu-rep
