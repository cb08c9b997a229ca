#!/bin/bash
# quick A/B: selected GPU tests, then the C2 bench line with and without the fused gradient
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "${TL_TESTS:-fused or host_api or loss_small or upstream or zero_cost}" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_quick.log
grep -v Warning gpurun_out/pytest_gpu_quick.log | tail -3
for v in ${TL_VARIANTS:-0 1}; do
  TL_NO_FUSED_GRAD=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fg$v.json 2> gpurun_out/bench_fg$v.err; echo "bench NO_FUSED_GRAD=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_fg$v.json')); print(d['ms_per_step'], d['roofline']['stage_ms']['persistence'], d['roofline']['stage_ms']['grad_fill_scatter'], d['e2e']['ms_per_step'])"
done
