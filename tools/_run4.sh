mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "golden_kernel_variants or per_bit or block_stream or stock_vdecode or frame_decode_equals" > gpurun_out/pytest_t32.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_t32.log
timeout 300 python tools/probe_tile32.py > gpurun_out/probe_tile32.log 2>&1; cat gpurun_out/probe_tile32.log
timeout 300 python tools/time_perbit.py > gpurun_out/time_perbit.log 2>&1; cat gpurun_out/time_perbit.log
cat > /tmp/single.py <<'PY'
import sys; sys.path.insert(0,'.')
import isee3_decoder_b200 as v
d=v.Viterbi224(4096); s=v.streams.telemetry_stream(4096,3.0,seed=1)[1]; p=d.dev_alloc(8192); d.h2d(p,s)
d.set_option("force_single",1); d.init(0); d.update_dev(p,4096)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_acs_single -s 100 -c 1 -o gpurun_out/r02_k_acs_single python /tmp/single.py > gpurun_out/ncu_single.log 2>&1; tail -3 gpurun_out/ncu_single.log
