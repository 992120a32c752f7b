mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "per_bit or block_stream or stock_vdecode or golden_default or recycled or empty_and or short_ring" > gpurun_out/pytest_pb.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_pb.log
timeout 300 python tools/time_perbit.py > gpurun_out/time_perbit.log 2>&1; cat gpurun_out/time_perbit.log
timeout 600 bash tools/time_dropin.sh > gpurun_out/time_dropin.log 2>&1; cat gpurun_out/time_dropin.log
timeout 300 python tools/time_startup.py > gpurun_out/time_startup.log 2>&1; cat gpurun_out/time_startup.log
timeout 300 python tools/time_frames.py > gpurun_out/time_frames.log 2>&1; tail -5 gpurun_out/time_frames.log
