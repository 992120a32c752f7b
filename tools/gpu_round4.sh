#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_fused.py 2048 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -c 1 -o gpurun_out/prof_persist python tools/prof_fused.py 2048 > gpurun_out/ncu.log 2>&1
cat gpurun_out/prof_plain.log; tail -5 gpurun_out/ncu.log
