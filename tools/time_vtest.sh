#!/bin/bash
# Wall clock of the reference's own vtest224.c (BASELINE config 1: -l 10000 -e 3) linked against libviterbi224_b200, two trial counts so
# that process start / CUDA initialisation can be separated from the per-frame cost (init + update(10000) + chainback(10000), vtest224.c:116-118).
cd "$(dirname "$0")/.."
python3 - <<'PY'
import subprocess, time
def run(n):
    t = time.time()
    out = subprocess.run(["oracle/_ref/vtest224_b200", "-l", "10000", "-e", "3", "-n", str(n)], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    return time.time() - t, out
run(1)
a, oa = run(36)
b, ob = run(236)
per = (b - a) / 200
print(f"vtest224_b200 -l 10000 -e 3 -n 36 : {a:.2f} s wall | {oa}")
print(f"vtest224_b200 -l 10000 -e 3 -n 236: {b:.2f} s wall | {ob}")
print(f"per 10,000-bit frame (init + update_blk + chainback through the stock program): {1e3 * per:.2f} ms = {10000 / per / 1e3:.0f} kbit/s; "
      f"process start and CUDA initialisation: {a - 36 * per:.2f} s")
PY
