#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python tools/prof_fused.py 8192 > gpurun_out/prof_plain.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/smoke.log; tail -8 gpurun_out/pytest_gpu.log; tail -c 300 gpurun_out/bench.err; cat gpurun_out/prof_plain.log
