mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "golden or per_bit or block_stream or stock_vdecode or frame_decode_equals or empty_and or recycled" > gpurun_out/pytest_fast2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_fast2.log
timeout 300 python tools/time_perbit.py > gpurun_out/time_perbit.log 2>&1; cat gpurun_out/time_perbit.log
timeout 600 bash tools/time_dropin.sh > gpurun_out/time_dropin.log 2>&1; head -2 gpurun_out/time_dropin.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], json.dumps(d['roofline_other_kernels']))"
