"""Short driver for ncu: a few hundred fused passes on a mid-stream state (no warm-up effects of init)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isee3_decoder_b200 as v224

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bits, syms = v224.streams.telemetry_stream(n, 3.0, seed=5)
with v224.Viterbi224(n) as d:
    d.init(0)
    d.kernel_time_enable(True)
    d.update_blk(syms, n)
    ms, k, passes = d.kernel_time_ms()
    print(f"{k} launches, {passes} passes, {1e3 * ms / passes:.2f} us per pass", d.stats())
