#!/bin/bash
# full GPU suite on the default library, then phase cycles + bench line per library variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | tail -4
for lib in "$@"; do
  echo "== $lib"
  TL_PROBE_LIB=$lib timeout 200 python scripts/r2_probe.py 2>&1 | head -2 | cut -c1-330
  TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stage_ms']['persistence'])"
done
