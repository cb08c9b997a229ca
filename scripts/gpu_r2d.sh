#!/bin/bash
# Fused gradient (tl_forward_backward) check: GPU suite, then the C2 bench line with and without the fusion, C5 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | tail -6
for v in 0 1; do
  TL_NO_FUSED_GRAD=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fg$v.json 2> gpurun_out/bench_fg$v.err; echo "bench NO_FUSED_GRAD=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_fg$v.json')); print(d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['ms_per_step'], d['clocks'])"
done
timeout 300 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print(d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['ms_per_step'])"
