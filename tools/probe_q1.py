"""A decoder alone on the 32-column build with ONE register per row and thread (256 threads per tile; option tile32 = 3) against the
two-register build (tile32 = 1) and the 64-column build (tile32 = 0); lockstep decoders on it (knob tile32 = 4).
(needs a library built with -DV224_WITH_Q1: tools/build_variants.sh q1:"-DV224_WITH_Q1")
usage: probe_q1.py [variant|default] [nbits]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import isee3_decoder_b200 as v224
from isee3_decoder_b200 import binding
name = sys.argv[1] if len(sys.argv) > 1 else "default"
if name != "default":
    binding.library_path = lambda: os.path.join(ROOT, "tools", "_bin", f"libv224_{name}.so")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
decs = [v224.Viterbi224(n) for _ in range(3)]
dptr = []
for i, d in enumerate(decs):
    s = v224.streams.telemetry_stream(n, 3.0, seed=50 + i)[1]
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
d = decs[0]
ref = None
per_sm = 4 if name == "q1c4" else 3
for tile32, grids in ((0, (148,)), (1, (444,)), (3, sorted({148, 222, 296, 370, 444, 148 * per_sm}))):
    for grid in grids:
        best = None
        for rep in range(3):
            d.init(0)
            d.set_option("tile32", tile32)
            d.set_option("grid_limit", grid)
            d.kernel_time_enable(True)
            d.update_dev(dptr[0], n)
            ms, k, passes = d.kernel_time_ms()
            us = 1e3 * ms / passes
            best = us if best is None or us < best else best
        m = d.get_metrics()
        if ref is None:
            ref = m
        print(f"{name}: alone, tile32 {tile32} grid {grid}: {best:.2f} us per pass   metrics identical: {bool(np.array_equal(m, ref))}", flush=True)
for nctx in (2, 3):
    for grid in (0,):
        best = None
        for rep in range(3):
            for x in decs[:nctx]:
                x.init(0)
            d.set_option("tile32", 4)
            d.set_option("grid_limit", grid)
            d.kernel_time_enable(True)
            v224.Viterbi224.update_multi_dev(decs[:nctx], dptr[:nctx], n)
            ms, k, passes = d.kernel_time_ms()
            us = 1e3 * ms / passes
            best = us if best is None or us < best else best
        print(f"{name}: {nctx} decoders in lockstep on the one-register build: {best:.2f} us per pass per decoder", flush=True)
