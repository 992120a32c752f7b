"""Short driver for ncu: N decoders in lockstep, a few hundred fused passes each.  usage: prof_multi.py <variant|default> <nctx> <nbits>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
from isee3_decoder_b200 import binding
name = sys.argv[1] if len(sys.argv) > 1 else "default"
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
if name != "default":
    binding.library_path = lambda: os.path.join(ROOT, "tools", "_bin", f"libv224_{name}.so")
decs = [v224.Viterbi224(n) for _ in range(nctx)]
dptr = []
for i, d in enumerate(decs):
    s = v224.streams.telemetry_stream(n, 3.0, seed=50 + i)[1]
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
for rep in range(2):
    for d in decs:
        d.init(0)
    decs[0].kernel_time_enable(True)
    v224.Viterbi224.update_multi_dev(decs, dptr, n)
    ms, k, passes = decs[0].kernel_time_ms()
print(f"{name}: {nctx} decoders, {passes} passes, {1e3 * ms / passes:.2f} us per pass per decoder")
