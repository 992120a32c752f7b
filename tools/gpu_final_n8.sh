#!/bin/bash
# round-end evidence on an 8-GPU box: the strong-scaling bench under torchrun and the same decode in one process
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; tail -c 300 gpurun_out/bench_n8.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n8.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','scaling')}, d['e2e']['value'], d['check'])"
timeout 600 python bench.py --native-multi 8 --steps 2 --warmup 1 > gpurun_out/bench_native8.log 2> gpurun_out/bench_native8.err; echo "native rc=$?"; tail -c 300 gpurun_out/bench_native8.err; cut -c1-700 gpurun_out/bench_native8.log
