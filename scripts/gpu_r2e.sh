#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | tail -4
timeout 200 python scripts/r2_probe.py > gpurun_out/r2_probe.log 2>&1; head -2 gpurun_out/r2_probe.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['ms_per_step'])"
TL_NO_FUSED_MATCH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nofuse.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_nofuse.json')); print('nofuse', d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['ms_per_step'])"
