#!/bin/bash
# final check of the round: whole GPU suite, smoke, bench line (C2) and the C5 line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | grep -E "passed|failed|FAILED|ERROR|rc=" | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 120 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
python - <<'P'
import json
for f in ("gpurun_out/r2_bench_c2_n1.json", "gpurun_out/r2_bench_c5_n1.json"):
    d = json.loads(open(f).read())
    print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"]["samples"], d["roofline"]["frac"])
P
