"""Small workload for compute-sanitizer (memcheck): batched frames side by side (concurrent tracebacks), a short segmented
stream decode, the per-bit ABI pattern.  Sizes are tiny because the sanitizer slows kernels by one to two orders of magnitude:
    compute-sanitizer --tool memcheck python tools/memcheck_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import isee3_decoder_b200 as v224
S = v224.streams
rng = np.random.default_rng(1)
fb, nframes = 64, 6
bits = rng.integers(0, 2, fb * nframes, dtype=np.uint8)
sym01, _ = S.encode_bits(bits, 0)
syms = S.awgn_symdemod(sym01, 5.0, rng)
with v224.Viterbi224(fb) as d:
    out = d.decode_frames(syms, nframes, fb, None, None, 4)
    d.init(0); d.update_blk(syms[: 2 * fb], fb); one = d.chainback(fb, 0)
    assert np.array_equal(out[0], one)
_, soft = S.telemetry_stream(3 * 4096, 4.0, seed=2)
with v224.Viterbi224(64 + 512) as d:
    d.init(0)
    a, _ = d.stream_decode(soft, 64)
    d.init(0)
    for i in range(40):
        d.update_blk(soft[2 * i: 2 * i + 2], 1); d.decodebit(32, 0)
print("memcheck workload done")
