#!/bin/bash
# A/B with parity check of the variant first
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pairs or kats or golden or loss_256 or c5" 2>&1 | tail -1
  TL_PROBE_LIB=$lib timeout 200 python scripts/r2_probe.py 2>&1 | head -2 | cut -c1-330
  TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stage_ms']['persistence'])"
done
