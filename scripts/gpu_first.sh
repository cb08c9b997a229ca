#!/bin/bash
# first GPU contact: parity tests + a timing probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python scripts/time_probe.py > gpurun_out/time_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/time_probe.log
cat gpurun_out/time_probe.log
