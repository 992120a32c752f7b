#!/bin/bash
# bench + ncu launch list of the same command + one full capture of the dominant kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_short.log 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/prof_fused.py 2048 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -c 1 -o gpurun_out/prof_persist python tools/prof_fused.py 2048 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/bench_short.err; wc -l gpurun_out/launches.csv; tail -2 gpurun_out/ncu_full.log
