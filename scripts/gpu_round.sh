#!/bin/bash
# tests + bench + ncu launch list (shares, not absolutes)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ph_|seg_sort|match_kernel|loss_kernel|grad_kernel' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
timeout 120 python scripts/stats_probe.py > gpurun_out/stats_probe.log 2>&1; grep -E 'pred dim1|truth dim1|iid dim1' gpurun_out/stats_probe.log
timeout 120 python scripts/probe_callsite.py > gpurun_out/callsite.json 2> gpurun_out/callsite.err; cat gpurun_out/callsite.json; tail -2 gpurun_out/callsite.err
