mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "golden or per_bit or block_stream or stock_vdecode or frame_decode_equals or empty_and" > gpurun_out/pytest_fast1.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_fast1.log
timeout 300 python tools/time_perbit.py > gpurun_out/time_perbit.log 2>&1; cat gpurun_out/time_perbit.log
for v in default diag t32c4 t32c6 slot3; do timeout 200 python tools/probe_tile32.py $v; done > gpurun_out/probe_tile32_ab.log 2>&1; cat gpurun_out/probe_tile32_ab.log
timeout 600 bash tools/time_dropin.sh > gpurun_out/time_dropin.log 2>&1; head -2 gpurun_out/time_dropin.log
