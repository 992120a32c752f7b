mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -k "config2 or range_decode or multi_gpu or short_ring or hybridtest or golden_default or segmented or native_block" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_new.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.log
timeout 300 python tools/probe_grid.py > gpurun_out/probe_grid.log 2>&1; cat gpurun_out/probe_grid.log
