mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "config2 or multi_gpu or short_ring or hybridtest" > gpurun_out/pytest_new2.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_new2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --total-bits 2097152 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/bench_n2.err; cut -c1-3000 gpurun_out/bench_n2.log
timeout 600 python bench.py --native-multi 2 --steps 2 --warmup 1 --total-bits 2097152 > gpurun_out/bench_native2.log 2> gpurun_out/bench_native2.err; echo "native rc=$?"; tail -c 800 gpurun_out/bench_native2.err; cat gpurun_out/bench_native2.log
