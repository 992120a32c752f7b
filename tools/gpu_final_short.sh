#!/bin/bash
# evidence refresh after a change of the fused pass only: smoke, GPU tests, bench, ncu launch list of the bench command, full captures
# of the persistent kernel (4 decoders in lockstep; one alone).  tools/refresh_profiles.py copies the results into profiles/.
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-rooflines > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 120 python tools/prof_multi.py default 4 2048 > gpurun_out/prof_multi_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -s 1 -c 1 -f -o gpurun_out/r02_k_acs_persist_4dec python tools/prof_multi.py default 4 2048 > gpurun_out/ncu_full.log 2>&1
timeout 120 python tools/prof_single.py alone 2048 > gpurun_out/prof_alone_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -s 1 -c 1 -f -o gpurun_out/r02_k_acs_persist_alone_t32 python tools/prof_single.py alone 2048 > gpurun_out/ncu_full_alone.log 2>&1
timeout 100 python tools/time_frames.py > gpurun_out/time_frames.log 2>&1; tail -4 gpurun_out/time_frames.log
cat gpurun_out/prof_multi_plain.log gpurun_out/prof_alone_plain.log; tail -2 gpurun_out/ncu_full.log; wc -l gpurun_out/launches.csv
cut -c1-300 gpurun_out/bench.log
