#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resample.py tests/test_dice_ce.py tests/test_training_step.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v Warning gpurun_out/pytest_gpu.log | tail -6
timeout 300 python scripts/probe_callsite.py > gpurun_out/callsite.json 2> gpurun_out/callsite.err; cat gpurun_out/callsite.json; tail -3 gpurun_out/callsite.err
