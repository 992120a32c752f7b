#!/bin/bash
# round-end evidence: tests, bench, ncu launch list of the same bench command, one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 400 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 2 --warmup 3"
timeout 600 $CMD > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"


wc -l gpurun_out/launches.csv
cat gpurun_out/bench.log
