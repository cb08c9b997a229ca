#!/bin/bash
# rectangular-map parity + the whole GPU suite (no -x: every failure is listed), then the bench line (NVML clock sampler)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | grep -E "passed|failed|FAILED|ERROR|rc=" | head -40
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2_bench_c2_n1.json").read())
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["roofline"]["frac"])
P
