#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/trace_passes.py 1 > gpurun_out/trace_1dec.log 2>&1
timeout 120 python tools/trace_passes.py 2 > gpurun_out/trace_2dec.log 2>&1
cat gpurun_out/trace_1dec.log | head -60; echo ======; head -40 gpurun_out/trace_2dec.log
