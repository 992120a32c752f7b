"""Frame-mode throughput (decode.c:220-222 pattern: init / update(1024) / chainback(1024) per frame): per-frame ABI calls vs
v224x_decode_frames with 1 and 3 frames side by side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import isee3_decoder_b200 as v224
S = v224.streams
nframes, fb = 48, 1024
rng = np.random.default_rng(7)
bits = S.telemetry_bits(nframes, rng)
sym01, _ = S.encode_bits(bits, 0)
syms = S.awgn_symdemod(sym01, 3.0, rng)
with v224.Viterbi224(fb) as d:
    for rep in range(2):
        t0 = time.perf_counter()
        for f in range(nframes):
            d.init(0); d.update_blk(syms[2 * fb * f: 2 * fb * (f + 1)], fb); d.chainback(fb, 0)
        t1 = time.perf_counter()
    print(f"per-frame ABI calls : {nframes / (t1 - t0):7.1f} frames/s  ({nframes * fb / (t1 - t0) / 1e3:.0f} kbit/s)")
    for lock in (1, 2, 3, 4):
        for rep in range(2):
            t0 = time.perf_counter(); d.decode_frames(syms, nframes, fb, None, None, lock); t1 = time.perf_counter()
        print(f"decode_frames lock {lock}: {nframes / (t1 - t0):7.1f} frames/s  ({nframes * fb / (t1 - t0) / 1e3:.0f} kbit/s)")
