mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/time_frames.py > gpurun_out/time_frames.log 2>&1; tail -5 gpurun_out/time_frames.log
