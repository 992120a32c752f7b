#!/bin/bash
# round-end evidence: tests, bench, ncu launch list of the same bench command, full captures of the dominant kernels
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-rooflines > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 120 python tools/prof_multi.py default 4 2048 > gpurun_out/prof_multi_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -s 1 -c 1 -o gpurun_out/r02_k_acs_persist_4dec python tools/prof_multi.py default 4 2048 > gpurun_out/ncu_full.log 2>&1
timeout 120 python tools/prof_single.py alone 2048 > gpurun_out/prof_alone_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_acs_persist -s 1 -c 1 -o gpurun_out/r02_k_acs_persist_alone_t32 python tools/prof_single.py alone 2048 > gpurun_out/ncu_full_alone.log 2>&1
timeout 120 python tools/prof_single.py stage 512 > gpurun_out/prof_stage_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_acs_single_fast -s 100 -c 1 -o gpurun_out/r02_k_acs_single_fast python tools/prof_single.py stage 512 > gpurun_out/ncu_full_stage.log 2>&1
timeout 300 python tools/report_chainback_redo.py 3 > gpurun_out/chainback_redo.log 2>&1; cat gpurun_out/chainback_redo.log
timeout 200 python tools/time_decode_block.py 1024 > gpurun_out/time_decode_block.log 2>&1; cat gpurun_out/time_decode_block.log
cat gpurun_out/prof_multi_plain.log gpurun_out/prof_alone_plain.log gpurun_out/prof_stage_plain.log; tail -2 gpurun_out/ncu_full.log; wc -l gpurun_out/launches.csv
cut -c1-400 gpurun_out/bench.log
