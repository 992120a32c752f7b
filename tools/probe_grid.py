"""Per-pass time of the persistent ACS kernel against the number of CTAs in the grid (option "grid_limit"), for 1..3 decoders in
lockstep.  A single decoder is latency bound (pass n+1 needs all of pass n): fewer, faster CTAs shorten the tile latency.
usage: probe_grid.py [nbits]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
decs = [v224.Viterbi224(n) for _ in range(3)]
dptr = []
for i, d in enumerate(decs):
    s = v224.streams.telemetry_stream(n, 3.0, seed=50 + i)[1]
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
for nctx in (1, 2, 3):
    for grid in (0, 148, 222, 296, 333, 370, 407):
        best = None
        for rep in range(3):
            for d in decs[:nctx]:
                d.init(0)
            decs[0].set_option("grid_limit", grid)
            decs[0].kernel_time_enable(True)
            v224.Viterbi224.update_multi_dev(decs[:nctx], dptr[:nctx], n)
            ms, k, passes = decs[0].kernel_time_ms()
            us = 1e3 * ms / passes
            best = us if best is None or us < best else best
        print(f"decoders {nctx} grid_limit {grid or 444}: {best:.2f} us per pass per decoder", flush=True)
decs[0].set_option("grid_limit", 0)
