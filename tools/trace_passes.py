"""Phase timeline of the persistent ACS kernel (needs tools/_bin/libv224_trace.so, tools/build_trace.sh).
usage: trace_passes.py [nctx]   -- events per (pass, tile) of decoder 0: 0 item claimed, 1 tile handed to the compute warps,
2 loads landed, 3 exchange read, 4 round 2 done, 5 stores + statistics issued, 6 done-word atom returned, 7 after resolve."""
import ctypes, os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
from isee3_decoder_b200 import binding
binding.library_path = lambda: os.path.join(ROOT, "tools", "_bin", "libv224_trace.so")
lib = v224.load_library()
nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = 64 * 8
decs = [v224.Viterbi224(n) for _ in range(nctx)]
dptr = []
for i, d in enumerate(decs):
    s = v224.streams.telemetry_stream(n, 3.0, seed=5 + i)[1]
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
for rep in range(2):
    for d in decs:
        d.init(0)
    v224.Viterbi224.update_multi_dev(decs, dptr, n)
tr = np.zeros(64 * 1024 * 8, dtype=np.uint64)
lib.v224_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
assert lib.v224_debug_read_trace(tr.ctypes.data_as(ctypes.c_void_p), tr.size) == 0
smid = np.zeros(64 * 1024, dtype=np.uint32)
lib.v224_debug_read_smid.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
assert lib.v224_debug_read_smid(smid.ctypes.data_as(ctypes.c_void_p), smid.size) == 0
smid = smid.reshape(64, 1024)[:, :512]
tr = tr.reshape(64, 1024, 8).astype(np.int64)[:, :512]
names = ["claimed", "handed over", "loads landed", "exchange read", "round2 done", "stores issued", "atom returned", "after resolve"]
print(f"{nctx} decoder(s) in lockstep (the trace holds the tiles of whichever decoder wrote last); times in us relative to the pass's first hand-over")
for p in range(20, 26):
    t = tr[p]
    base = t[:, 1].min()
    print(f"pass {p}: first hand-over -> last atom {1e-3 * (t[:, 6].max() - base):6.2f} us;  next pass first hand-over at {1e-3 * (tr[p + 1][:, 1].min() - base):6.2f}, "
          f"pass-after-next at {1e-3 * (tr[p + 2][:, 1].min() - base):6.2f}")
    for e in range(8):
        v = 1e-3 * (t[:, e] - base)
        print(f"    {names[e]:14s} min {v.min():7.2f}  p10 {np.percentile(v, 10):7.2f}  median {np.median(v):7.2f}  p90 {np.percentile(v, 90):7.2f}  max {v.max():7.2f}")
    ph = [np.median(t[:, i + 1] - t[:, i]) * 1e-3 for i in range(7)]
    print("    phase medians: claim->handover %.2f  load %.2f  round1+xchg %.2f  round2 %.2f  stores+stats %.2f  signal %.2f  resolve %.2f" % tuple(ph))
    dur = 1e-3 * (t[:, 5] - t[:, 1])
    print(f"    tile hand-over -> stores issued: min {dur.min():.2f} median {np.median(dur):.2f} p90 {np.percentile(dur, 90):.2f} max {dur.max():.2f}")
p = 23
t = tr[p]
by = collections.defaultdict(list)
for tile in range(512):
    by[int(smid[p, tile])].append(tile)
hist = collections.Counter(len(v) for v in by.values())
print(f"pass {p}: SMs by number of tiles of this pass:", dict(sorted(hist.items())), " SMs used:", len(by))
dur = 1e-3 * (t[:, 5] - t[:, 1])
for k in sorted(hist):
    sel = [tile for v in by.values() if len(v) == k for tile in v]
    print(f"  {k} tiles on the SM: hand-over -> stores issued median {np.median(dur[sel]):.2f} max {dur[sel].max():.2f}")
