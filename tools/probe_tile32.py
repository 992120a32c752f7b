"""Per-pass time of a decoder running alone: the 64-column-tile and the 32-column-tile build of the fused pass against the
number of CTAs in the grid.  usage: probe_tile32.py [variant|default] [nbits]   (variants: tools/build_variants.sh)"""
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
d = v224.Viterbi224(n)
s = v224.streams.telemetry_stream(n, 3.0, seed=50)[1]
p = d.dev_alloc(2 * n); d.h2d(p, s)
ref = None
per_sm = {"t32c4": 4, "t32c6": 6}.get(name, 5)
grids32 = sorted({148, 296, 444, 148 * per_sm, 148 * per_sm - 74})
for tile32, grids in ((0, (148,)), (1, grids32)):
    for grid in grids:
        best = None
        for rep in range(3):
            d.init(0)
            d.set_option("tile32", tile32)
            d.set_option("grid_limit", grid)
            d.kernel_time_enable(True)
            d.update_dev(p, n)
            ms, k, passes = d.kernel_time_ms()
            us = 1e3 * ms / passes
            best = us if best is None or us < best else best
        m = d.get_metrics()
        if ref is None:
            ref = m
        print(f"{name}: tile32 {tile32} grid {grid}: {best:.2f} us per pass   metrics identical to the first variant: {bool(np.array_equal(m, ref))}", flush=True)
