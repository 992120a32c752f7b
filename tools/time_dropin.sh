#!/bin/bash
# Wall-clock of the reference's own programs linked against libviterbi224_b200 (per-bit ABI use) vs the reference decoder.
# Process start + CUDA context creation (~2 s) is separated from the per-bit cost by timing two input lengths.
python - <<'PY'
import sys, time, subprocess, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import isee3_decoder_b200 as v
bits, soft = v.streams.telemetry_stream(65536, 4.0, seed=3)
def run(exe, n):
    p = os.path.join('oracle', '_ref', exe)
    t = time.time(); out = subprocess.run([p, '-d', '200', '-q'], input=soft[:2 * n].tobytes(), capture_output=True); dt = time.time() - t
    return dt, out
run('vdecode_b200', 1024)                                   # cold start of the box
t1 = min(run('vdecode_b200', 8192)[0] for _ in range(3))
t2, o2 = min((run('vdecode_b200', 65536) for _ in range(2)), key=lambda x: x[0])
print(f"vdecode_b200 (stock vdecode.c, update(1)+decodebit(200,0) per bit): 8192 pairs {t1:.2f} s, 65536 pairs {t2:.2f} s (best of 3 / 2) -> "
      f"{1e6 * (t2 - t1) / 57344:.0f} us per bit = {57344 / (t2 - t1):.0f} bits/s steady state, {t1 - 8192 * (t2 - t1) / 57344:.1f} s start-up (rc {o2.returncode}, {len(o2.stdout)} chars)")
(s1, _), (s2, o) = run('vdecode_sse', 256), run('vdecode_sse', 1024)
print(f"vdecode_sse  (reference SSE2 decoder, one host core):              256 pairs {s1:.2f} s, 1024 pairs {s2:.2f} s -> {768 / (s2 - s1):.0f} bits/s steady state")
for exe, args in (('vtest224_b200', ['-l', '8192', '-n', '4', '-e', '3']), ('vtest224sse', ['-l', '1024', '-n', '1', '-e', '3'])):
    p = os.path.join('oracle', '_ref', exe)
    t = time.time(); out = subprocess.run([p] + args, capture_output=True, text=True); dt = time.time() - t
    nb = int(args[1]) * int(args[3])
    print(f"{exe} {' '.join(args)}: {dt:.2f} s wall incl. process start  | {out.stdout.strip().splitlines()[-1]}")
PY
