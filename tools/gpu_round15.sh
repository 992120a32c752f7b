#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_multi.py default 3 2048 > gpurun_out/prof_multi_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -s 1 -c 1 -o gpurun_out/prof_ws_3dec_bulk python tools/prof_multi.py default 3 2048 > gpurun_out/ncu_multi.log 2>&1
cat gpurun_out/prof_multi_plain.log; tail -3 gpurun_out/ncu_multi.log
