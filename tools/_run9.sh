mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "golden_kernel_variants or frame_decode_equals" > gpurun_out/pytest_q1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_q1.log
for v in default q1c4; do timeout 200 python tools/probe_q1.py $v; done > gpurun_out/probe_q1.log 2>&1; cat gpurun_out/probe_q1.log
