"""Per-CTA / per-SM timeline of the persistent ACS kernel from the trace build (tools/build_trace.sh).
usage: trace_sm.py [nctx]   Events per tile: 0 claimed, 1 handed over (protocol), 2 input in registers, 3 exchange read,
4 round 2 done, 5 stores + statistics issued, 6 done-word atom returned, 7 after resolve."""
import ctypes, os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224
from isee3_decoder_b200 import binding
binding.library_path = lambda: os.path.join(ROOT, "tools", "_bin", "libv224_trace.so")
lib = v224.load_library()
nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 60 * 8
decs = [v224.Viterbi224(n) for _ in range(nctx)]
dptr = []
for i, d in enumerate(decs):
    s = v224.streams.telemetry_stream(n, 3.0, seed=5 + i)[1]
    p = d.dev_alloc(2 * n); d.h2d(p, s); dptr.append(p)
for rep in range(2):
    for d in decs:
        d.init(0)
    v224.Viterbi224.update_multi_dev(decs, dptr, n)
tr = np.zeros(64 * 4 * 512 * 8, dtype=np.uint64)
lib.v224_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
assert lib.v224_debug_read_trace(tr.ctypes.data_as(ctypes.c_void_p), tr.size) == 0
sm = np.zeros(64 * 4 * 512, dtype=np.uint32)
lib.v224_debug_read_smid.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
assert lib.v224_debug_read_smid(sm.ctypes.data_as(ctypes.c_void_p), sm.size) == 0
tr = tr.reshape(64, 4, 512, 8).astype(np.int64)
sm = sm.reshape(64, 4, 512)
P0, P1 = 15, 50                                  # steady-state passes
T = tr[P0:P1, :nctx].reshape(-1, 8)
S = sm[P0:P1, :nctx].reshape(-1)
smid, cta = S & 0xffff, S >> 16
t0 = T[:, 1].min()
T = (T - t0) * 1e-3                              # us
span = T[:, 5].max() - T[:, 2].min()
print(f"{nctx} decoder(s), passes {P0}..{P1 - 1}: {T.shape[0]} tiles in {span:.1f} us -> {span / (P1 - P0) / nctx:.2f} us per pass per decoder")
ph = lambda a, b: np.median(T[:, b] - T[:, a])
print(f"tile medians (us): claim->handover {ph(0,1):.2f}  handover->input in regs {ph(1,2):.2f}  round1+xchg {ph(2,3):.2f}  round2 {ph(3,4):.2f}  "
      f"stores+stats {ph(4,5):.2f}  ->atom returned {ph(5,6):.2f};  input->stores p10/50/90: "
      + "/".join(f"{np.percentile(T[:,5]-T[:,2], q):.2f}" for q in (10, 50, 90)))
# per CTA: gap between a tile's stores and the next tile's input
gaps = []
by_cta = collections.defaultdict(list)
for i in range(T.shape[0]):
    by_cta[int(cta[i])].append(i)
for c, idx in by_cta.items():
    idx.sort(key=lambda i: T[i, 2])
    for a, b in zip(idx[:-1], idx[1:]):
        gaps.append(T[b, 2] - T[a, 5])
gaps = np.array(gaps)
busy = np.array([sum(T[i, 5] - T[i, 2] for i in idx) for idx in by_cta.values()])
print(f"CTAs {len(by_cta)}: tiles per CTA {T.shape[0] / len(by_cta):.1f}; per-CTA busy fraction (input->stores) median {np.median(busy) / span:.2f}; "
      f"gap stores->next input p10/50/90/mean: " + "/".join(f"{np.percentile(gaps, q):.2f}" for q in (10, 50, 90)) + f"/{gaps.mean():.2f} us")
# per SM: how many of its CTAs are in a compute phase (input in registers .. round 2 done) at a time
lo, hi = np.percentile(T[:, 2], 5), np.percentile(T[:, 4], 95)
grid = np.arange(lo, hi, 0.05)
hist = np.zeros(5)
by_sm = collections.defaultdict(list)
for i in range(T.shape[0]):
    by_sm[int(smid[i])].append(i)
for s_, idx in by_sm.items():
    cnt = np.zeros(grid.size, dtype=np.int32)
    for i in idx:
        cnt += ((grid >= T[i, 2]) & (grid < T[i, 4])).astype(np.int32)
    for k in range(5):
        hist[k] += (cnt == k).sum()
hist /= hist.sum()
print("per SM, fraction of time with k CTAs in a compute round: " + "  ".join(f"k={k}: {hist[k]:.2f}" for k in range(4)) + f"   mean {sum(k * hist[k] for k in range(5)):.2f}")
tiles_per_sm = np.array([len(v) for v in by_sm.values()])
print(f"SMs {len(by_sm)}: tiles per SM min/median/max {tiles_per_sm.min()}/{int(np.median(tiles_per_sm))}/{tiles_per_sm.max()}")
