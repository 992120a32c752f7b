mkdir -p gpurun_out
timeout 600 python bench.py --config 5 --total-bits 4194304 --steps 1 --warmup 3 --no-cpu-baseline --no-side-rooflines > gpurun_out/bench_c5_smoke.log 2> gpurun_out/bench_c5_smoke.err; echo "c5 smoke rc=$?"; tail -c 600 gpurun_out/bench_c5_smoke.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_c5_smoke.log') if l.startswith('{')][-1])
print('c5', d['value'], d['e2e']['value'], d['check'])"
timeout 1200 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-side-rooflines > gpurun_out/bench_c4.log 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; tail -c 600 gpurun_out/bench_c4.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_c4.log') if l.startswith('{')][-1])
print('c4', d['value'], d['e2e']['value'], d['roofline']['frac'], d['check'])"
