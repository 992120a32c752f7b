"""bench.py's BER check on a rank whose stream has false-alarm phase flips (rank 6 of the 8-GPU workload)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import isee3_decoder_b200 as v224
for rank in (6, 1):
    wl = bench.make_workload(rank, 8)
    n = wl["npairs"]
    with v224.Viterbi224(bench.BLOCK + bench.DELAY) as dec:
        dec.init_uniform(5000, -1)
        out, rep = dec.stream_decode_seg(wl["pairs"], bench.DELAY, 3, bench.CONV)
    print(rank, "flips", wl["flips"], "-> errors, compared, transient:", bench.ber_check(out, wl), rep)
