"""Per-call cost of the per-bit ABI pattern (vdecode.c:145-152) on the GPU library."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import isee3_decoder_b200 as v224
n, delay = 2000, 200
bits, syms = v224.streams.telemetry_stream(n, 3.0, seed=43)
with v224.Viterbi224(delay + 1) as d:
    d.init(0)
    pairs = [np.ascontiguousarray(syms[2 * i: 2 * i + 2]) for i in range(n)]
    for i in range(200):
        d.update_blk(pairs[i], 1); d.decodebit(delay, 0)
    t0 = time.perf_counter()
    for i in range(200, 1100):
        d.update_blk(pairs[i], 1)
    t1 = time.perf_counter()
    for i in range(900):
        d.decodebit(delay, 0)
    t2 = time.perf_counter()
    for i in range(1100, 2000):
        d.update_blk(pairs[i], 1); d.decodebit(delay, 0)
    t3 = time.perf_counter()
    d.set_option("no_walk_cache", 1)
    for i in range(900):
        d.decodebit(delay, 0)
    t4 = time.perf_counter()
    print(f"update(1): {1e6 * (t1 - t0) / 900:.1f} us   decodebit (cached, same head): {1e6 * (t2 - t1) / 900:.1f} us   "
          f"update+decodebit: {1e6 * (t3 - t2) / 900:.1f} us -> {900 / (t3 - t2):.0f} bits/s   decodebit full walk: {1e6 * (t4 - t3) / 900:.1f} us   "
          f"walk steps {d.stats()['walk_steps']}")
