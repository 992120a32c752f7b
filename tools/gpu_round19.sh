#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 200 -k "golden or lockstep or segmented" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_variants.py default early latenobulk nobulk > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log
