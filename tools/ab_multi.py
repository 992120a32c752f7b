"""Throughput of 1..4 independent decoders advanced in lockstep by one persistent launch (GPU)."""
import os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isee3_decoder_b200 as v224
n = 16384
decs = [v224.Viterbi224(n) for _ in range(4)]
syms = [v224.streams.telemetry_stream(n, 3.0, seed=50 + i)[1] for i in range(4)]
dptr = []
for d, s in zip(decs, syms):
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
# reference results: each decoder alone
ref = []
for d, p in zip(decs, dptr):
    d.init(0); r = d.update_dev(p, n); ref.append((r, d.get_metrics().copy(), d.stats()["renormals"]))
for nctx in (1, 2, 3, 4):
    times = []
    for rep in range(4):
        for d in decs[:nctx]:
            d.init(0)
        decs[0].kernel_time_enable(True)
        ren = v224.Viterbi224.update_multi_dev(decs[:nctx], dptr[:nctx], n)
        ms, k, passes = decs[0].kernel_time_ms()
        if rep: times.append(1e3 * ms / passes)
    ok = all(ren[i] == ref[i][0] and np.array_equal(decs[i].get_metrics(), ref[i][1]) and decs[i].stats()["renormals"] == ref[i][2] for i in range(nctx))
    t = statistics.median(times)
    print(f"{nctx} decoders: {t:6.2f} us per pass per decoder -> aggregate {8e3 / t:7.1f} kbit/s   identical to separate runs: {ok}")
