"""A/B timing of the ACS kernel variants on one stream (GPU): interleaved repetitions, median / min."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isee3_decoder_b200 as v224
n = 16384
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
bits, syms = v224.streams.telemetry_stream(n, 3.0, seed=5)
variants = [("dynamic queue", {"tile_mode": 0}), ("static coop", {"tile_mode": 1}), ("balanced", {"tile_mode": 2}),
            ("per-pass launch", {"per_pass_launch": 1}), ("single-stage", {"force_single": 1})]
res = {name: [] for name, _ in variants}
with v224.Viterbi224(n) as d:
    for r in range(reps + 1):
        for name, opts in variants:
            for k in ("tile_mode", "per_pass_launch", "force_single"):
                d.set_option(k, -1 if k == "tile_mode" else 0)
            for k, v in opts.items():
                d.set_option(k, v)
            nn = n if "single" not in name else 1024
            d.init(0); d.kernel_time_enable(True); d.update_blk(syms, nn)
            ms, k, passes = d.kernel_time_ms()
            if r > 0:
                res[name].append(1e3 * ms / (passes if passes else nn))
for name, _ in variants:
    v = res[name]
    print(f"{name:18s} median {statistics.median(v):7.2f}  min {min(v):7.2f}  max {max(v):7.2f} us per {'stage' if 'single' in name else 'pass'}  ({len(v)} reps)")
