"""Communication floor of the state-space-sharded variant, measured: with the 2^23 path metrics sharded over G GPUs
by the top log2(G) state bits, every fused pass of k = 8 stages ends in a perfect-shuffle all-to-all of the whole 16 MiB
metric array ((G-1)/G of every GPU's 16/G MiB shard leaves the GPU).  This times exactly that exchange (NCCL all_to_all_single
over NVLink, uint16 payload of the real size, device events, max over ranks) -- the compute of a pass is NOT included, so the
figure is a lower bound on the sharded variant's pass time, to be set against the time-segmented variant's measured pass time
(bench.py: ~9.3 us per pass per GPU, each GPU working on its own segment with no exchange at all).

    python -m torch.distributed.run --nproc-per-node G tools/sharded_exchange_bench.py
"""
import json
import os
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shard = (1 << 23) // world                       # states per GPU
send = torch.randint(0, 256, (2 * shard,), dtype=torch.uint8, device="cuda")   # the shard as bytes (NCCL has no 16-bit integer type)
recv = torch.empty_like(send)
iters, warm = 200, 20
for _ in range(warm):
    dist.all_to_all_single(recv, send)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    dist.all_to_all_single(recv, send)
e1.record(); torch.cuda.synchronize()
us = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], device="cuda")
dist.all_reduce(us, op=dist.ReduceOp.MAX)
# CUDA-graph variant: 50 exchanges per launch (what a fused kernel could at best approach without its own P2P protocol)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        dist.all_to_all_single(recv, send)
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g, stream=s):
            for _ in range(50):
                dist.all_to_all_single(recv, send)
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0.record(s)
        for _ in range(4):
            g.replay()
        e1.record(s); torch.cuda.synchronize()
        gus = torch.tensor([e0.elapsed_time(e1) * 1e3 / 200], device="cuda")
    except Exception as ex:                         # graph capture of NCCL may be unavailable
        gus = torch.tensor([float("nan")], device="cuda")
dist.all_reduce(gus, op=dist.ReduceOp.MAX)
torch.cuda.synchronize()
if rank == 0:
    out = world - 1
    print(json.dumps({"gpus": world, "metric_bytes_total": 2 << 23, "bytes_leaving_each_gpu_per_pass": 2 * shard * out // world,
                      "all_to_all_us_per_pass": float(us), "all_to_all_us_per_pass_cuda_graph": float(gus),
                      "stages_per_exchange": 8,
                      "floor_bits_per_s_whole_job": 8 / (float(min(us, gus) if gus == gus else us) * 1e-6),
                      "note": "exchange only, no ACS compute; time-segmented variant: 8 / 9.3 us = 0.86 Mbit/s PER GPU, no exchange"}), flush=True)
os._exit(0)      # (destroy_process_group after an NCCL graph capture can hang; nothing left to clean up)
