mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --config 5 --total-bits 67108864 --steps 1 --warmup 3 > gpurun_out/bench_c5_n8.log 2> gpurun_out/bench_c5_n8.err; echo "c5 n8 rc=$?"; tail -c 500 gpurun_out/bench_c5_n8.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_c5_n8.log') if l.startswith('{')][-1])
print('c5 n8', d['value'], d['e2e']['value'], d['ms_per_step'], d['check'])"
