"""Throughput of the 32-column-tile build with lockstep decoders (measurement knob tile32 = 2): how much of a lone decoder's pass time
is dependency stall, how much the smaller tile's own overhead.  usage: probe_t32_multi.py [nbits]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
decs = [v224.Viterbi224(n) for _ in range(4)]
dptr = []
for i, d in enumerate(decs):
    s = v224.streams.telemetry_stream(n, 3.0, seed=50 + i)[1]
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
for tile32, grids in ((0, (0,)), (2, (444, 592, 740))):
    for nctx in (1, 2, 3, 4):
        for grid in grids:
            best = None
            for rep in range(3):
                for d in decs[:nctx]:
                    d.init(0)
                decs[0].set_option("tile32", tile32)
                decs[0].set_option("grid_limit", grid)
                decs[0].kernel_time_enable(True)
                v224.Viterbi224.update_multi_dev(decs[:nctx], dptr[:nctx], n)
                ms, k, passes = decs[0].kernel_time_ms()
                us = 1e3 * ms / passes
                best = us if best is None or us < best else best
            print(f"tile cols {64 if tile32 == 0 else 32} decoders {nctx} grid {grid or 'default'}: {best:.2f} us per pass per decoder", flush=True)
