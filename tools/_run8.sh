mkdir -p gpurun_out
run() { # name lib segments
  V224_LIB=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-rooflines --segments $3 > gpurun_out/ab_$1.log 2> gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/ab_{n}.log') if l.startswith('{')][-1])
    print(f"{n:16s} value {d['value']/1e3:7.1f} kbit/s  e2e {d['e2e']['value']/1e3:7.1f}  pass {d['roofline']['mean_pass_us']:.2f} us  frac {d['roofline']['frac']:.3f}  clocks {d['clocks']['sm_mhz']} MHz  power {d['clocks'].get('power_w_max')} W  reasons {d['clocks']['reasons']}  residual {d['check']['residual_diffs_vs_one_gpu_decode']}")
except Exception as e:
    print(n, "FAILED", e, open(f'gpurun_out/ab_{n}.err').read()[-300:])
PY
}
run base3 "" 3
run base4 "" 4
run s200_3 tools/_bin/libv224_s200.so 3
run s200_4 tools/_bin/libv224_s200.so 4
run s800_4 tools/_bin/libv224_s800.so 4
for t in 1 4; do V224_PAIR_THREADS=$t python - <<'PY'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, isee3_decoder_b200 as v
_, soft = v.streams.telemetry_stream(8*1024*1024, 3.0, seed=9, junk_symbols=101)
ts=[]
for i in range(4):
    t=time.perf_counter(); a, fa = v.pair_symbols(soft); ts.append(time.perf_counter()-t)
print("pair threads", os.environ["V224_PAIR_THREADS"], "16M symbols:", " ".join(f"{x*1e3:.1f}" for x in ts), "ms", os.cpu_count(), len(os.sched_getaffinity(0)))
PY
done
