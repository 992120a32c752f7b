#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 200 -k "golden or lockstep or segmented or empty" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_variants.py default nobulk > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log
timeout 100 python tools/trace_sm.py 3 2>&1 | tail -6
