#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err; cat gpurun_out/bench.log
