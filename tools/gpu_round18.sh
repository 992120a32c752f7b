#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 -k "native_block or stock_vdecode" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, time, subprocess, os
sys.path.insert(0, '.')
import isee3_decoder_b200 as v
bits, soft = v.streams.telemetry_stream(1 << 20, 3.0, seed=3, junk_symbols=101)
blk = os.path.join('isee3-decoder_b200', 'bin', 'vdecode_block')
for n in (1 << 18, 1 << 20):
    t = time.time(); out = subprocess.run([blk, '-d', '200', '-q'], input=soft[:2 * n].tobytes(), capture_output=True); dt = time.time() - t
    print(f"vdecode_block: {n} pairs in {dt:.2f} s wall incl. process start (rc {out.returncode}, {len(out.stdout)} chars)")
PY
