#!/usr/bin/env python
"""BASELINE config 5: box-scale throughput sweep -- one long synthetic stream (default 268,435,456 bits, 512 MiB of soft
symbols), time-segmented over the GPUs of the box.

    python tools/config5.py [--bits N] [--ebn0 dB]                                         (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/config5.py

Strong scaling: the stream is fixed, rank g decodes the contiguous range [g*N/G, (g+1)*N/G) with a leading warm-up of
2048 stages from uniform metrics (isee3-decoder_b200/segments.py), and inside the rank 3 decoders advance in lockstep
over contiguous sub-segments with device-verified hand-overs (v224x_stream_decode_seg_dev).  No data-path collective.

The stream is generated ON the GPU (torch Philox: data bits, the K=24 encoder of encode.c:17-35 as shifted XORs,
symdemod-format AWGN quantisation of symdemod.c:240-251) -- parity is judged on identical symbol bytes, not identical
random numbers -- and every decoded bit is compared with the transmitted data on the GPU (BER).  The first
1,048,576 stages of rank 0 are also decoded sequentially with the plain single-decoder path and compared bit for bit.

Prints one JSON line (rank 0).  The state-sharded variant is not run here: its exchange step alone was measured at
23-32 us per 8-stage pass (profiles/r01_sharded_exchange_floor.jsonl), i.e. <= 0.34 Mbit/s for the whole box."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DELAY = 200
BLOCK = 8192
WARM = 2048
CONV = 2048
SEGMENTS = 3
SEED = 50505
GEN_BLOCK = 1 << 24     # the stream's generation unit (bits)
K = 24
POLY1 = 0o73665667
POLY2 = 0o73665665
LAG = DELAY + K - 2          # output i of the streaming decode is data bit i - LAG (vdecode.c:152,176-177)


def gen_bits(torch, dev, seed, index, n):
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1000 + index)
    return torch.randint(0, 2, (n,), dtype=torch.uint8, device=dev, generator=g)


def encode(torch, hist23, bits):
    """encode.c:17-35: register bit i at time t is d[t - i] (bit 0 newest); symbols parity(reg & POLY1), !parity(reg & POLY2)."""
    d = torch.cat([hist23, bits])
    n = bits.numel()
    s1 = torch.zeros(n, dtype=torch.uint8, device=bits.device)
    s2 = torch.ones(n, dtype=torch.uint8, device=bits.device)          # G2FLIP (code.h:63)
    for i in range(K):
        seg = d[K - 1 - i: K - 1 - i + n]
        if (POLY1 >> i) & 1:
            s1 ^= seg
        if (POLY2 >> i) & 1:
            s2 ^= seg
    return s1, s2


def soften(torch, s1, s2, ebn0_db, seed, index, chunk=1 << 24):
    """symdemod wire format: total RMS 100 around 128, clipped to [0, 255], truncated (symdemod.c:190,240-251)."""
    esn0 = 10 ** (ebn0_db / 10.0) * 0.5
    sigma = 100.0 / (1.0 + 2.0 * esn0) ** 0.5
    amp = sigma * (2.0 * esn0) ** 0.5
    n = s1.numel()
    soft = torch.empty(2 * n, dtype=torch.uint8, device=s1.device)
    g = torch.Generator(device=s1.device)
    g.manual_seed(seed * 1000 + 500 + index)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        for k, s in enumerate((s1, s2)):
            y = (2.0 * s[a:b].float() - 1.0) * amp + sigma * torch.randn(b - a, device=s1.device, generator=g) + 128.0
            soft[2 * a + k: 2 * b: 2] = y.clamp_(0, 255).to(torch.uint8)
    return soft


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=1 << 28)
    ap.add_argument("--ebn0", type=float, default=3.0)
    ap.add_argument("--segments", type=int, default=SEGMENTS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import isee3_decoder_b200 as v224
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                        # NCCL's banner must not land on stdout
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    per = args.bits // world
    assert per * world == args.bits and per % GEN_BLOCK == 0, f"bits per GPU must be a multiple of {GEN_BLOCK}"
    # The stream is defined in blocks of GEN_BLOCK bits keyed by (seed, block index) -- data and noise -- so that its bytes
    # do not depend on the number of GPUs: runs at different N decode the SAME stream and their outputs can be compared.
    t_gen = time.perf_counter()

    def block(j):
        bits_j = gen_bits(torch, dev, SEED, j, GEN_BLOCK)
        hist = gen_bits(torch, dev, SEED, j - 1, GEN_BLOCK)[-(K - 1):].clone() if j > 0 else torch.zeros(K - 1, dtype=torch.uint8, device=dev)
        s1, s2 = encode(torch, hist, bits_j)
        return bits_j, soften(torch, s1, s2, args.ebn0, SEED, j)

    first = rank * per // GEN_BLOCK
    parts = [block(j) for j in range(first, first + per // GEN_BLOCK)]
    if rank == 0:
        lead_bits = torch.zeros(0, dtype=torch.uint8, device=dev)
        lead_soft = torch.zeros(0, dtype=torch.uint8, device=dev)
    else:
        pb, ps = block(first - 1)
        lead_bits, lead_soft = pb[-WARM:].clone(), ps[-2 * WARM:].clone()
        del pb, ps
    lead = lead_bits
    data = torch.cat([lead_bits] + [p[0] for p in parts])
    soft = torch.cat([lead_soft] + [p[1] for p in parts])
    del parts
    n = data.numel()
    out = torch.zeros(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    dec = v224.Viterbi224(BLOCK + DELAY, device=local_rank)

    def start():
        dec.init(0) if rank == 0 else dec.init_uniform(5000, -1)

    # warm-up (clocks, allocations of the lockstep decoders) + sequential cross-check on a prefix
    npre = min(n, 1 << 20)
    start()
    dec.stream_decode_dev(soft.data_ptr(), npre, DELAY, out.data_ptr())
    seq = out[:npre].clone()
    start()
    dec.stream_decode_seg_dev(soft.data_ptr(), npre, DELAY, out.data_ptr(), args.segments, CONV)
    prefix_same = bool(torch.equal(seq, out[:npre]))
    out.zero_()

    l0 = dec.stats()
    barrier()
    t0 = time.perf_counter()
    dec.timer_start()
    start()
    rep = dec.stream_decode_seg_dev(soft.data_ptr(), n, DELAY, out.data_ptr(), args.segments, CONV)
    ms = dec.timer_stop_ms()
    barrier()
    wall = time.perf_counter() - t0
    st = dec.stats()

    # BER on the GPU: outputs [skip, n) of this rank are data bits [skip - LAG, n - LAG) of `data`
    skip = lead.numel() if rank else LAG
    got = out[skip:]
    want = data[skip - LAG: n - LAG]
    wrong = torch.nonzero(got != want).flatten()
    errs = int(wrong.numel())
    checked = int(got.numel())
    # absolute data-bit positions of the errors (at most 512 per rank), gathered on rank 0
    pos = torch.full((512,), -1, dtype=torch.int64, device=dev)
    k = min(512, errs)
    pos[:k] = wrong[:k] + (skip - LAG) + (rank * per - lead.numel())
    near_start = int((wrong < 8192).sum().item()) if rank else 0           # errors right behind a rank's warm-up

    t = torch.tensor([ms, wall * 1e3, t_gen * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(errs), float(checked), float(st["launches"] - l0["launches"]), float(rep.get("redone", 0)),
                        float(rep.get("extra_stages", 0) + (lead.numel())), float(near_start)], dtype=torch.float64, device=dev)
    allpos = [pos]
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        allpos = [torch.zeros_like(pos) for _ in range(world)]
        dist.all_gather(allpos, pos)
    if rank == 0:
        ms_max, wall_max, gen_max = (float(x) for x in t)
        errs_t, checked_t, launches_t, redone_t, extra_t, near_t = (float(x) for x in tot)
        positions = sorted(int(x) for x in torch.cat(allpos).tolist() if x >= 0)
        value = args.bits / (ms_max * 1e-3)
        print(json.dumps({
            "config": "BASELINE config 5: box-scale throughput sweep, time-segmented",
            "bits": args.bits, "n_gpus": world, "bits_per_gpu": per, "ebn0_db": args.ebn0, "decode_delay": DELAY, "scaling": "strong",
            "format": "symdemod-format soft symbols generated on the GPU (Philox), pairs fed directly (no phase search)",
            "decoded_bits_per_s": value, "state_updates_per_s": value * (1 << 23), "device_ms": ms_max, "wall_ms": wall_max,
            "generation_ms": gen_max, "warmup_stages_per_rank": WARM, "segments_per_gpu": rep.get("segments"),
            "handovers_verified_rank0": rep.get("verified"), "segments_redone_all_ranks": int(redone_t),
            "extra_stages_all_ranks": int(extra_t), "overhead_frac": extra_t / args.bits,
            "bit_errors_vs_transmitted": int(errs_t), "bits_checked": int(checked_t), "ber": errs_t / max(1.0, checked_t),
            "bit_errors_within_8192_bits_of_a_rank_start": int(near_t), "error_positions": positions,
            "prefix_1M_segmented_identical_to_sequential_rank0": prefix_same, "gpu_launches": int(launches_t),
            "passes_rank0": {k: st[k] - l0[k] for k in ("fused_passes", "careful_passes", "single_stages", "sat_stages")},
            "state_sharded_variant": "not kept: exchange floor 23.4 us/pass at G=8, 31.7 us at G=2 (profiles/r01_sharded_exchange_floor.jsonl)"}),
            flush=True)
    dec.delete()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
