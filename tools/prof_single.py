"""Short drivers for ncu: 'alone' = one decoder alone in the persistent kernel (32-column-tile build), 'stage' = the one-stage
kernel back to back.  usage: prof_single.py alone|stage [nbits]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
d = v224.Viterbi224(n)
s = v224.streams.telemetry_stream(n, 3.0, seed=50)[1]
p = d.dev_alloc(2 * n); d.h2d(p, s)
if mode == "stage":
    d.set_option("force_single", 1)
for rep in range(2):
    d.init(0)
    d.kernel_time_enable(True)
    d.update_dev(p, n)
    ms, k, passes = d.kernel_time_ms()
print(f"{mode}: {n} stages, {1e3 * ms / max(1, passes if mode == 'alone' else n):.2f} us per {'pass' if mode == 'alone' else 'stage'}")
