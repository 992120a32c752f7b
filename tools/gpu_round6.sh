#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_variants.py default nq2c4 nq2c3 nq4c4 nq4c3 > gpurun_out/ab_variants.log 2>&1
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/ab_variants.log
