"""Phase timeline of the persistent ACS kernel (needs tools/_bin/libv224_trace.so built with -DV224_TRACE)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
from isee3_decoder_b200 import binding
binding.library_path = lambda: os.path.join(ROOT, "tools", "_bin", "libv224_trace.so")
lib = v224.load_library()
n = 64 * 8
bits, syms = v224.streams.telemetry_stream(n, 3.0, seed=5)
static = int(sys.argv[1]) if len(sys.argv) > 1 else 0
with v224.Viterbi224(n) as d:
    d.init(0); d.update_blk(syms, n)
    d.init(0); d.update_blk(syms, n)
    tr = np.zeros(64 * 1024 * 8, dtype=np.uint64)
    lib.v224_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
    rc = lib.v224_debug_read_trace(tr.ctypes.data_as(ctypes.c_void_p), tr.size)
    assert rc == 0, rc
    smid = np.zeros(64 * 1024, dtype=np.uint32)
    lib.v224_debug_read_smid.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
    assert lib.v224_debug_read_smid(smid.ctypes.data_as(ctypes.c_void_p), smid.size) == 0
smid = smid.reshape(64, 1024)
tr = tr.reshape(64, 1024, 8).astype(np.int64)[:, :512]
smid = smid[:, :512]
names = ["claim", "dep ok", "loads landed", "exchange done", "round2 done", "stores issued", "fence done", "signalled"]
print("static" if static else "dynamic", "tiles; times in us relative to the pass's first 'dep ok'")
for p in range(20, 28):
    t = tr[p]
    base = t[:, 1].min()
    print(f"pass {p}: length (first dep ok -> last signalled) {1e-3 * (t[:, 7].max() - base):6.2f} us;  next pass first dep ok at {1e-3 * (tr[p + 1][:, 1].min() - base):6.2f}")
    for e in range(8):
        v = 1e-3 * (t[:, e] - base)
        print(f"    {names[e]:14s} min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f}")
    dur = 1e-3 * (t[:, 7] - t[:, 1])
    print(f"    per-tile dep ok -> signalled: min {dur.min():.2f} median {np.median(dur):.2f} max {dur.max():.2f};  phases median: load {np.median(t[:,2]-t[:,1])*1e-3:.2f}  round1 {np.median(t[:,3]-t[:,2])*1e-3:.2f}  round2 {np.median(t[:,4]-t[:,3])*1e-3:.2f}  stats+store {np.median(t[:,5]-t[:,4])*1e-3:.2f}  fence {np.median(t[:,6]-t[:,5])*1e-3:.2f}  signal {np.median(t[:,7]-t[:,6])*1e-3:.2f}")

print("per-SM view of pass 25: active tiles on the SM -> round-1 duration (us) of those tiles")
p = 25
t = tr[p]
r1 = 1e-3 * (t[:, 3] - t[:, 2])
tot = 1e-3 * (t[:, 6] - t[:, 1])
import collections
by = collections.defaultdict(list)
for tile in range(512):
    by[int(smid[p, tile])].append(tile)
hist = collections.Counter(len(v) for v in by.values())
print("  SMs by number of active tiles:", dict(sorted(hist.items())), " SMs used:", len(by))
for k in sorted(hist):
    sel = [tile for v in by.values() if len(v) == k for tile in v]
    print(f"  {k} active: round1 median {np.median(r1[sel]):.2f} max {r1[sel].max():.2f};  dep ok->fence median {np.median(tot[sel]):.2f} max {tot[sel].max():.2f}")
worst = np.argsort(-tot)[:8]
for tile in worst:
    sm = int(smid[p, tile]); print(f"  slow tile {tile}: sm {sm} with {len(by[sm])} active tiles, round1 {r1[tile]:.2f}, total {tot[tile]:.2f}")
