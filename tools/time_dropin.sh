#!/bin/bash
# Wall-clock of the reference's own programs linked against libviterbi224_b200 (per-bit ABI use) vs the reference decoder.
python - <<'PY'
import sys, time, subprocess, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import isee3_decoder_b200 as v, pyoracle
bits, soft = v.streams.telemetry_stream(4096, 4.0, seed=3)
for exe in ('vdecode_b200', 'vdecode_sse'):
    p = os.path.join('oracle', '_ref', exe)
    n = 4096 if exe.endswith('b200') else 1024
    t = time.time(); out = subprocess.run([p, '-d', '200', '-q'], input=soft[:2 * n].tobytes(), capture_output=True); dt = time.time() - t
    print(f"{exe}: {n} pairs in {dt:.2f} s -> {n / dt:.0f} bits/s  (rc {out.returncode}, {len(out.stdout)} chars)")
for exe, args in (('vtest224_b200', ['-l', '8192', '-n', '4', '-e', '3']), ('vtest224sse', ['-l', '1024', '-n', '1', '-e', '3'])):
    p = os.path.join('oracle', '_ref', exe)
    t = time.time(); out = subprocess.run([p] + args, capture_output=True, text=True); dt = time.time() - t
    nb = int(args[1]) * int(args[3])
    print(f"{exe} {' '.join(args)}: {dt:.2f} s wall -> {nb / dt:.0f} bits/s incl. process start  | {out.stdout.strip().splitlines()[-1]}")
PY
