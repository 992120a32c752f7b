#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_variants.py nq2c4o1 nq2c3o1 nq2c3 nq4c4o1 nq4c3o1 nq4c3 > gpurun_out/ab_variants.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/ab_variants.log
