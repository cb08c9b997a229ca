#!/bin/bash
# parity tests + phase shares (quick iteration loop); every step under its own timeout
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 90 python scripts/stats_probe.py > gpurun_out/stats_probe.log 2>&1; grep -E 'dim1' gpurun_out/stats_probe.log | grep -v smooth3
timeout 90 python scripts/time_probe.py > gpurun_out/time_probe.log 2>&1; grep -E "^B |pairs-only|phase" gpurun_out/time_probe.log
