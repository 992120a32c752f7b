"""A/B timing of the ACS kernel variants on one stream (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isee3_decoder_b200 as v224
n = 16384
bits, syms = v224.streams.telemetry_stream(n, 3.0, seed=5)
for name, opts in [("balanced", {"tile_mode": 2}), ("dynamic queue", {"tile_mode": 0}), ("static coop", {"tile_mode": 1}), ("per-pass launch", {"per_pass_launch": 1}), ("single-stage", {"force_single": 1})]:
    with v224.Viterbi224(n) as d:
        for k, v in opts.items():
            d.set_option(k, v)
        nn = n if "single" not in name else 2048
        d.init(0); d.update_blk(syms, nn)          # warm
        d.init(0); d.kernel_time_enable(True); d.update_blk(syms, nn)
        ms, k, passes = d.kernel_time_ms()
        per = 1e3 * ms / (passes if passes else nn)
        print(f"{name:18s} {per:8.2f} us per {'pass' if passes else 'stage'}   ({k} launches)")
