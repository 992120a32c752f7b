mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_discard.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_discard.log
run() { # name segments extra
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side-rooflines --segments $2 > gpurun_out/ab_$1.log 2> gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/ab_{n}.log') if l.startswith('{')][-1])
    print(f"{n:16s} value {d['value']/1e3:7.1f} kbit/s  e2e {d['e2e']['value']/1e3:7.1f}  pass {d['roofline']['mean_pass_us']:.2f} us  frac {d['roofline']['frac']:.3f}  clocks {d['clocks']['sm_mhz']} MHz  power {d['clocks'].get('power_w_max')} W  reasons {d['clocks']['reasons']}  residual {d['check']['residual_diffs_vs_one_gpu_decode']} errs {d['check']['bit_errors_vs_transmitted']} careful {d['check']['passes']}")
except Exception as e:
    print(n, "FAILED", e, open(f'gpurun_out/ab_{n}.err').read()[-300:])
PY
}
run seg3 3
run seg4 4
run seg4b 4
timeout 300 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:k_acs_persist -s 1 -c 1 python tools/prof_multi.py default 4 2048 2>&1 | grep -E "dram__|gpu__time|us per pass"
