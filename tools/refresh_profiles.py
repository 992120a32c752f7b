"""Copies the evidence of the last tools/gpu_final.sh run (gpurun_out/) into profiles/ under the round's names: bench lines, ncu launch
list + per-kernel summary, ncu --set full details pages, the DRAM bytes per pass of the dominant kernel (profiles/fused_traffic.json).
usage: python tools/refresh_profiles.py [round tag, default r02]"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


for src, dst in (("bench.log", f"{tag}_bench_n1.json"), ("bench_ref.log", f"{tag}_bench_reference_arm.json")):
    d = last_json(os.path.join(G, src))
    json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
    print(dst, d["value"], (d.get("roofline") or {}).get("frac"), d.get("clocks"))
for src, dst in (("launches.csv", f"{tag}_launches_bench.csv"), ("pytest_gpu.log", f"{tag}_pytest_gpu.txt"), ("chainback_redo.log", f"{tag}_chainback_redo.txt"),
                 ("time_decode_block.log", f"{tag}_time_decode_block.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
# per-kernel shares of the launch list
rows = list(csv.reader(l for l in open(os.path.join(P, f"{tag}_launches_bench.csv")) if not l.startswith("==")))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except (ValueError, IndexError):
        continue
    us = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3}.get(r[ui], v)
    k = r[ki].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-rooflines",
       "(first 400 launches of the run, final library: 4 decoders in lockstep; per-launch times under ncu are cold-cache and serialised -- compare SHARES, not absolutes)", "",
       f"{'kernel':28s} {'launches':>8s} {'total us':>12s} {'share':>8s} {'mean us':>10s}"]
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"{k:28s} {n:8d} {t:12.1f} {100 * t / tot:7.2f}% {t / n:10.1f}")
open(os.path.join(P, f"{tag}_launches_bench_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[3:]))
# full captures
for rep in ("k_acs_persist_4dec", "k_acs_persist_alone_t32", "k_acs_single_fast"):
    src = os.path.join(G, f"{tag}_{rep}.ncu-rep")
    if os.path.exists(src):
        txt = subprocess.run(["ncu", "-i", src, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{tag}_ncu_full_{rep}.txt"), "w").write(txt)
src = os.path.join(G, f"{tag}_k_acs_persist_4dec.ncu-rep")
raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def g(k):
    i = hdr.index(k)
    return float(vals[i].replace(",", "")) * scale[units[i]]


passes = 1024
traffic = {"kernel": "k_acs_persist",
           "source": f"profiles/{tag}_ncu_full_k_acs_persist_4dec.txt (ncu --set full, one launch of 256 passes x 4 decoders in lockstep, final library)",
           "dram_bytes_read_per_pass": g("dram__bytes_read.sum") / passes, "dram_bytes_write_per_pass": g("dram__bytes_write.sum") / passes,
           "dram_bytes_per_pass": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / passes, "algorithmic_bytes_per_pass": 41943056,
           "us_per_pass_under_ncu": g("gpu__time_duration.sum") / passes,
           "note": "DRAM sees the decision rows (8 MiB per pass) plus what is left of the path-metric write-back: the retirer warp drops a tile's consumed input "
                   "lines from the L2 (discard.global.L2) for every pass that cannot be invalidated, so the dead lines of the 4 x 3 x 16 MiB of metric buffers are "
                   "no longer written back when they are evicted (round 1 / option no_discard: 25 MB per pass); metric reads hit L2"}
json.dump(traffic, open(os.path.join(P, "fused_traffic.json"), "w"), indent=1)
print({k: v for k, v in traffic.items() if k.startswith(("dram", "us_"))})
