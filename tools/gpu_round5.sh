#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_multi.py > gpurun_out/ab_multi.log 2>&1
timeout 300 python tools/trace_passes.py 0 > gpurun_out/trace_dyn.log 2>&1
cat gpurun_out/ab_multi.log; head -40 gpurun_out/trace_dyn.log
