"""A/B timing of library builds (tools/build_variants.sh) at the bench's shape: per-pass time of the persistent ACS kernel with
4 decoders in lockstep (64-column build) and with one decoder alone (32-column build), each variant in its own process, long
enough to reach the board's power steady state; final metrics CRC printed so that variants can be checked against each other.
usage: python tools/ab_variants4.py default m11 m55 ..."""
import os, sys, subprocess, statistics, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import isee3_decoder_b200 as v224
    from isee3_decoder_b200 import binding
    name = sys.argv[2]
    if name != "default":
        binding.library_path = lambda: os.path.join(ROOT, "tools", "_bin", f"libv224_{name}.so")
    n = 8192
    nmax = 4
    decs = [v224.Viterbi224(n) for _ in range(nmax)]
    syms = [v224.streams.telemetry_stream(n, 3.0, seed=50 + i)[1] for i in range(nmax)]
    dptr = []
    for d, s in zip(decs, syms):
        p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
    out = []
    for nctx, reps in ((4, 48), (1, 24)):
        times = []
        for rep in range(reps):
            if rep % 8 == 0:
                for d in decs[:nctx]:
                    d.init(0)
            decs[0].kernel_time_enable(True)
            v224.Viterbi224.update_multi_dev(decs[:nctx], dptr[:nctx], n)
            ms, k, passes = decs[0].kernel_time_ms()
            if rep >= reps // 3: times.append(1e3 * ms / passes)
        out.append((nctx, statistics.median(times), min(times)))
    crc = zlib.crc32(decs[0].get_metrics().tobytes())
    st = decs[0].stats()
    print(f"{name:10s} " + "  ".join(f"{nctx} dec: median {t:6.3f} min {m:6.3f} us/pass ({8e3 / t:6.0f} kbit/s)" for nctx, t, m in out)
          + f"   crc {crc:08x} careful {st['careful_passes']} renormals {st['renormals']}", flush=True)
else:
    for name in sys.argv[1:]:
        r = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or f"{name}: FAILED rc {r.returncode}\n{r.stderr[-1500:]}", flush=True)
