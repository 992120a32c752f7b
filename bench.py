#!/usr/bin/env python
"""bench.py -- decoded bits/s of the B200 viterbi224 decoder on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--total-bits B] [--scaling strong|weak] [--impl reference]

One "step" = one complete decode of the workload's soft-symbol stream through the streaming path of vdecode.c (the symbol
pairing / phase flip of vdecode.c:101-140 on the host, then update + decodebit(delay, state 0) per pair, vdecode.c:145-152)
in block form: fused 8-stage ACS passes + batched tracebacks, per GPU 4 decoders that one persistent kernel advances in
lockstep over contiguous segments (every hand-over verified on the device, output identical to the sequential decode).

  config 2 (default) : symdemod-format telemetry stream at 3 dB, decode delay 200, odd junk prefix (automatic phase flip).
                       N = 1: 1,048,576 bits (BASELINE's size).  N > 1: ONE fixed stream of 8,388,608 bits cut into N time
                       segments (strong scaling; --scaling weak: 1,048,576 bits per GPU, --total-bits to change the stream)
  config 3           : 4,194,304 bits, vtest-style AWGN at 2 dB, decode delay 2048 (long traceback, late survivor merge)
  config 4           : 16,777,216 bits, vtest-style AWGN at 1 dB (no frame structure: pairs as received, vdecode -F)
  config 5           : one long stream generated on the GPUs (N > 1: 268,435,456 bits; N = 1: 33,554,432 bits unless --total-bits
                       says otherwise; 2^24-bit blocks keyed by seed and block index, symdemod format at 3 dB), strong scaling

  value : whole-job decoded bits/s with the symbol pairs resident in HBM (device events, max over ranks); N > 1 includes
          the rank-to-rank hand-over verification (16 MiB metric snapshot per range over NCCL + the check kernel)
  e2e   : the same from HOST buffers through the C ABI: host pairing / phase flip (v224x_pair_symbols, once over the whole
          stream), H2D of the rank's pairs from pinned memory, decode, verification, gather of all decoded bits on rank 0
          and D2H -- all inside the timed region
  roofline : k_acs_persist (the dominant kernel) -- algorithmic bytes per launch / mean launch duration from CUDA events
          on the library's stream; plus the single-decoder persistent launch and the one-stage kernel k_acs_single (N = 1)
  cpu_baseline : the reference's own SSE2 decoder (oracle/_ref, compiled from the unmodified sources) on all host cores

N > 1: one process per GPU (torchrun); the pair stream is cut into N contiguous ranges with a warm-up of delay + 2048
stages per range; the only exchange is the hand-over check and the final gather (isee3-decoder_b200/segments.py).
The same decode inside ONE process (v224x_multi_stream_decode: one host thread per GPU, peer copies) is timed by
`--native-multi N` (not a torchrun mode; used for profiles/).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DELAY = 200                     # vdecode default decode delay (vdecode.c:44)
BLOCK = 8192                    # stages per update batch; ring = BLOCK + DELAY rows (8.2 GiB)
SEGMENTS = 4                    # per GPU: decoders advanced in lockstep over contiguous segments of the rank's range
CONV = 2048                     # stages a late-started decoder gets to converge before its verified hand-over
SEED = 20141
JUNK = 101                      # config 2: odd junk prefix, vdecode starts on the wrong symbol phase
FK = 8
K = 24
B_PASS = 2 * (1 << 24) + FK * (1 << 20) + 2 * FK      # algorithmic bytes of one fused pass
B_STAGE_UNFUSED = 34603010                             # bytes per decoded bit of the unfused algorithm (one k_acs_single launch)
CONFIGS = {
    2: {"bits_n1": 1 << 20, "bits_multi": 1 << 23, "ebn0": 3.0, "style": "symdemod telemetry (1024-bit minor frames, sync word), odd junk prefix",
        "pairing": "vdecode.c:101-140 sync correlator, automatic phase flip"},
    3: {"bits_n1": 1 << 22, "bits_multi": 1 << 22, "ebn0": 2.0, "delay": 2048, "style": "vtest-style AWGN, random data, long decode delay (late survivor merge)",
        "pairing": "as received (vdecode -F)"},
    4: {"bits_n1": 1 << 24, "bits_multi": 1 << 24, "ebn0": 1.0, "style": "vtest-style AWGN, random data", "pairing": "as received (vdecode -F)"},
    5: {"bits_n1": 1 << 25, "bits_multi": 1 << 28, "ebn0": 3.0, "style": "symdemod-format AWGN, random data, generated on the GPU",
        "pairing": "as received (vdecode -F)"},
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------------
def host_stream(config, total_bits):
    """The whole received stream on the host (configs 2 and 4): (transmitted bits, soft symbols)."""
    import isee3_decoder_b200 as v224
    S = v224.streams
    if config == 2:
        return S.telemetry_stream(total_bits, CONFIGS[2]["ebn0"], seed=SEED, junk_symbols=JUNK)
    rng = np.random.default_rng(SEED + config)
    bits = rng.integers(0, 2, total_bits, dtype=np.uint8)
    sym01, _ = S.encode_bits(bits, 0)
    return bits, S.awgn_vtest(sym01, CONFIGS[config]["ebn0"], rng)


GEN_BLOCK = 1 << 24     # config 5: the stream's generation unit (bits)
POLY1 = 0o73665667
POLY2 = 0o73665665


def gpu_stream_range(torch, dev, first, last):
    """Config 5: (data bits, soft symbols) of stream stages [first, last) as device tensors.  The stream is defined in
    2^24-bit blocks keyed by (seed, block index), so its bytes do not depend on how many GPUs decode it (encode.c:17-35 as
    shifted XORs over the bit history, symdemod.c:240-251 quantisation; parity is judged on identical bytes)."""
    import isee3_decoder_b200 as v224
    amp, sigma = v224.streams.symdemod_amplitudes(CONFIGS[5]["ebn0"])

    def block_bits(index):
        g = torch.Generator(device=dev)
        g.manual_seed(SEED * 1000 + index)
        return torch.randint(0, 2, (GEN_BLOCK,), dtype=torch.uint8, device=dev, generator=g)

    b0, b1 = first // GEN_BLOCK, (last - 1) // GEN_BLOCK
    bits_parts, sym_parts = [], []
    hist = block_bits(b0 - 1)[-(K - 1):] if b0 > 0 else torch.zeros(K - 1, dtype=torch.uint8, device=dev)
    for bi in range(b0, b1 + 1):
        bits = block_bits(bi)
        d = torch.cat([hist, bits])
        s1 = torch.zeros(GEN_BLOCK, dtype=torch.uint8, device=dev)
        s2 = torch.ones(GEN_BLOCK, dtype=torch.uint8, device=dev)          # G2FLIP (code.h:63)
        for i in range(K):
            seg = d[K - 1 - i: K - 1 - i + GEN_BLOCK]
            if (POLY1 >> i) & 1:
                s1 ^= seg
            if (POLY2 >> i) & 1:
                s2 ^= seg
        g = torch.Generator(device=dev)
        g.manual_seed(SEED * 1000 + 500000 + bi)
        sym01 = torch.stack([s1, s2], dim=1).reshape(-1).to(torch.float32)
        y = (2.0 * sym01 - 1.0) * amp + sigma * torch.randn(2 * GEN_BLOCK, device=dev, generator=g) + 128.0
        soft = y.clamp_(0, 255).to(torch.uint8)
        lo, hi = max(first, bi * GEN_BLOCK) - bi * GEN_BLOCK, min(last, (bi + 1) * GEN_BLOCK) - bi * GEN_BLOCK
        bits_parts.append(bits[lo:hi])
        sym_parts.append(soft[2 * lo: 2 * hi])
        hist = bits[-(K - 1):]
    return torch.cat(bits_parts), torch.cat(sym_parts)


def ber_check(out_bits, tx_bits, flips, delay, base_shift, first=0):
    """Decoded output vs transmitted data (lag = delay + K - 2 pairs, vdecode.c:176-177), window by window.

    vdecode's sync correlator (vdecode.c:107-140) flips the symbol phase whenever the out-of-phase correlation peak of a frame
    beats the in-phase one: by design behind the odd junk prefix, and -- at 3 dB, as in the reference -- now and then as a
    false alarm that the next frame corrects.  Every flip drops one symbol; between a false alarm and its correction the
    decoder is fed mis-paired symbols (garbage out, the reference prints the same), and after the correction the output is
    one more pair ahead.  Only windows within reach of a recorded flip (its frame, the next one, and the decoder's memory)
    are excused as transients; a window with many errors anywhere else counts in full, so burst errors of a broken
    kernel cannot hide here.  out_bits[i] belongs to pair first + i.  Returns (bit errors, bits compared, transient bits)."""
    n = out_bits.size
    lag = delay + K - 2
    win = 4096
    reach = 2048 + delay + 4096                        # pairs: a false alarm lasts one frame (2048 pairs), then the decoder re-converges
    flips = sorted(flips)
    # alignments: the junk prefix delays the data by base_shift pairs; a false alarm and its correction drop two symbols
    # (the output runs one pair further ahead)
    shifts = list(range(base_shift - len(flips) - 1, base_shift + 2))
    errs = compared = transient = 0
    for a in range(0, n, win):
        b = min(n, a + win)
        pa, pb = first + a, first + b                  # pair indices of the window
        near_flip = any(f - win <= pb and pa <= f + reach for f in flips)
        idx = np.arange(pa, pb)
        best = None
        for shift in shifts:
            src = idx - lag - shift
            ok = (src >= 0) & (src < tx_bits.size)
            if ok.sum() < idx.size // 2:
                continue
            e = int((out_bits[idx[ok] - first] != tx_bits[src[ok]]).sum())
            if best is None or e < best[0]:
                best = (e, int(ok.sum()))
        if best is None:
            continue
        if near_flip and best[0] > best[1] // 128:     # wrong symbol phase / re-acquisition right after a recorded flip
            transient += best[1]
        else:
            errs += best[0]
            compared += best[1]
    return errs, compared, transient


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------------------------
def run_reference_sample(sample_bits, threads, config=2):
    """The reference's SSE2 decoder on `threads` host cores, one independent decoder per core, each running
    the vdecode.c:145-152 loop (update(1) + decodebit(200, 0)) over `sample_bits` pairs.  Returns (bits/s, kind, s)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    import isee3_decoder_b200 as v224
    Dec = pyoracle.best_cpu_decoder()
    kind = "reference" if Dec is pyoracle.RefSSE2 else "port"
    _, soft = v224.streams.telemetry_stream(sample_bits, CONFIGS[config]["ebn0"], seed=SEED + 7)
    decs = [Dec(DELAY + 1) for _ in range(threads)]
    for d in decs:
        d.init(0)
    barrier = threading.Barrier(threads + 1)

    def work(d):
        barrier.wait()
        d.stream_decode(soft, DELAY, sample_bits)
        barrier.wait()

    ts = [threading.Thread(target=work, args=(d,)) for d in decs]
    for t in ts:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in ts:
        t.join()
    for d in decs:
        d.delete()
    return threads * sample_bits / dt, kind, dt


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def workload_config(args, world, total_bits):
    c = CONFIGS[args.config]
    return {"workload": f"BASELINE config {args.config}: vdecode streaming decode, {c['style']}, {total_bits} bits in all, Eb/N0 {c['ebn0']} dB, "
                        f"decode delay {DELAY}; pairing: {c['pairing']}",
            "total_bits": total_bits, "bits_per_gpu": total_bits // world, "decode_delay": DELAY, "stages_per_pass": FK, "block_stages": BLOCK,
            "parallelism": (f"ONE stream time-segmented over {world} GPUs (one process per GPU), warm-up {DELAY}+{CONV} stages per range, every "
                            f"rank-to-rank hand-over verified (16 MiB metric snapshot over NCCL + check kernel, exact redo on failure); "
                            if world > 1 else "single GPU; ")
                           + f"per GPU {args.segments} decoders in lockstep over contiguous segments, hand-overs verified on the device "
                             f"(warm-up {DELAY}+{CONV} stages each)",
            "cache": "the decision rings (8.2 GiB per decoder) are written once per stage and are far larger than the 126 MB L2; the "
                     "16 MiB path-metric buffers (3 per decoder) are re-read by the next pass by construction (no artificial L2 "
                     "flush possible without changing the algorithm)"}


def reference_arm(args, rank, world, total_bits):
    if rank != 0:
        return
    threads = host_threads()
    sample = 384
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, dt = run_reference_sample(sample, threads, args.config if args.config in (2,) else 2)
        if i >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    ms = 1e3 * statistics.mean(dt for _, dt in vals)
    sample_desc = (f"{threads} decoders (one per host core) x {sample} symbol pairs of the same symdemod-format stream, "
                   f"update(1)+decodebit(200,0) per bit; len {DELAY + 1} ring")
    line = {"impl": "reference", "metric": "decoded_bits_per_s", "value": value, "unit": "bits/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling_of(args, world), "vs_baseline": None, "dtype": "s16",
            "data": "synthetic", "config": workload_config(args, world, total_bits),
            "cpu_baseline": {"value": value, "unit": "bits/s", "cores": threads, "kind": kind, "sample": sample_desc},
            "e2e": {"value": value, "unit": "bits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "state_updates_per_s": value * (1 << 23), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def scaling_of(args, world):
    return "weak" if (world == 1 or args.scaling == "weak") else "strong"


def total_bits_of(args, world):
    if args.total_bits:
        return int(args.total_bits)
    c = CONFIGS[args.config]
    if world == 1:
        return c["bits_n1"]
    if args.scaling == "weak":
        return c["bits_n1"] * world if args.config == 2 else c["bits_multi"]
    return c["bits_multi"]


# ------------------------------------------------------------------------------------------------------------------
# the single-kernel rooflines beside the headline (N = 1): one-stage kernel and single-decoder persistent launch
# ------------------------------------------------------------------------------------------------------------------
def side_rooflines(v224, dec, dsyms_ptr, peak):
    out = {}
    nb = 4096
    # k_acs_single: the per-bit ABI pattern's kernel (vdecode.c:145), 34,603,010 algorithmic bytes per launch
    dec.set_option("force_single", 1)
    dec.init(0)
    dec.update_dev(dsyms_ptr, 256)
    dec.kernel_time_enable(True)
    dec.update_dev(dsyms_ptr, nb)
    ms, launches, _ = dec.kernel_time_ms()
    dec.kernel_time_enable(False)
    dec.set_option("force_single", 0)
    if ms > 0:
        ach = B_STAGE_UNFUSED * nb / (ms * 1e-3) / 1e9
        out["k_acs_single"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                               "algorithmic_bytes_per_launch": B_STAGE_UNFUSED, "mean_launch_us": 1e3 * ms / nb,
                               "note": f"{nb} back-to-back one-stage launches (events around the batch)"}
    # one decoder alone in the persistent kernel: what every stock frame caller (vtest224, decode.c, hybridtest) gets
    nb = 65536
    dec.init(0)
    dec.update_dev(dsyms_ptr, 8192)
    dec.kernel_time_enable(True)
    dec.update_dev(dsyms_ptr, nb)
    ms, launches, passes = dec.kernel_time_ms()
    dec.kernel_time_enable(False)
    if ms > 0 and passes:
        ach = B_PASS * passes / (ms * 1e-3) / 1e9
        out["k_acs_persist_single_decoder"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                               "algorithmic_bytes_per_pass": B_PASS, "mean_pass_us": 1e3 * ms / passes, "passes_timed": passes,
                                               "decoded_bits_per_s": nb / (ms * 1e-3)}
    return out


# ------------------------------------------------------------------------------------------------------------------
def native_multi(args):
    """One process, N GPUs through v224x_multi_stream_decode (one host thread per GPU inside the library)."""
    import isee3_decoder_b200 as v224
    n_gpu = args.native_multi
    total_bits = int(args.total_bits) if args.total_bits else CONFIGS[2]["bits_multi"]
    tx, soft = host_stream(2, total_bits)
    lib = v224.load_library()
    hp_soft = lib.v224x_host_alloc_pinned(soft.size)
    ctypes.memmove(hp_soft, soft.ctypes.data, soft.size)
    cap = soft.size // 2 + 1
    hp_pairs = lib.v224x_host_alloc_pinned(2 * cap)
    hp_bits = lib.v224x_host_alloc_pinned(cap)
    bits_view = np.ctypeslib.as_array(ctypes.cast(hp_bits, ctypes.POINTER(ctypes.c_uint8)), (cap,))
    flips = np.zeros(4096, np.int64)
    nf = ctypes.c_int(0)
    m = v224.MultiGpu(n_gpu, BLOCK + DELAY)
    rep = v224.binding.MultiReport()

    def step():
        npairs = lib.v224x_pair_symbols(hp_soft, soft.size, 0, 0, DELAY, hp_pairs, None, flips.ctypes.data_as(ctypes.c_void_p), flips.size, ctypes.byref(nf))
        m.init(0)
        rc = lib.v224x_multi_stream_decode(m.h, hp_pairs, npairs, DELAY, hp_bits, args.segments, CONV, ctypes.byref(rep))
        assert rc == 0, lib.v224x_last_error()
        return npairs

    for _ in range(max(1, args.warmup)):
        npairs = step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    out = bits_view[:npairs].copy()
    # the same pair stream on ONE GPU, sequential lockstep decode: must be identical bit for bit
    pairs = np.ctypeslib.as_array(ctypes.cast(hp_pairs, ctypes.POINTER(ctypes.c_uint8)), (2 * npairs,))
    with v224.Viterbi224(BLOCK + DELAY, device=0) as d:
        d.init(0)
        one, _ = d.stream_decode_seg(pairs, DELAY, args.segments, CONV)
    fl = [int(x) for x in flips[: nf.value]]
    errs, nchk, ntrans = ber_check(out, tx, fl, DELAY, (JUNK - 1) // 2)
    line = {"mode": "native-multi (one process, v224x_multi_stream_decode)", "metric": "decoded_bits_per_s", "value": npairs / dt, "unit": "bits/s",
            "n_gpus": n_gpu, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "timing": "host wall clock around the C call "
            "(pairing, H2D, decode, verification, D2H inside)", "total_bits": total_bits,
            "report": {k: getattr(rep, k) for k, _ in v224.binding.MultiReport._fields_},
            "check": {"residual_diffs_vs_one_gpu_decode": int((out != one).sum()), "bit_errors_vs_transmitted": errs, "bits_checked": nchk,
                      "bits_in_phase_flip_transients": ntrans, "phase_flips": fl}}
    print(json.dumps(line), flush=True)
    m.delete()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--total-bits", type=int, default=0, help="length of the stream (default: the config's size; N > 1: the fixed strong-scaling stream)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: one fixed stream (strong) or 1 Mi bits per GPU (weak)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-rooflines", action="store_true")
    ap.add_argument("--segments", type=int, default=SEGMENTS, help="decoders advanced in lockstep per GPU (1 = sequential)")
    ap.add_argument("--native-multi", type=int, default=0, help="one process, this many GPUs through v224x_multi_stream_decode")
    args = ap.parse_args()
    global DELAY
    DELAY = CONFIGS[args.config].get("delay", DELAY)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    total_bits = total_bits_of(args, world)

    if args.impl == "reference":
        reference_arm(args, rank, world, total_bits)
        return
    if args.native_multi:
        native_multi(args)
        return

    import isee3_decoder_b200 as v224
    import torch
    dist = None
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator is built; stdout carries exactly one JSON
        # line, so the file descriptor points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = v224.load_library()
    dec = v224.Viterbi224(BLOCK + DELAY, device=local_rank)
    nseg = max(1, args.segments)
    rd = v224.segments.GpuRangeDecoder(dec, torch, dev, nseg)
    W = DELAY + CONV

    # ---------------- the stream ----------------
    on_gpu = args.config == 5
    dontflip = args.config != 2
    if not on_gpu:
        tx_bits, soft = host_stream(args.config, total_bits)
        nsoft = soft.size
        hp_soft = lib.v224x_host_alloc_pinned(nsoft)
        ctypes.memmove(hp_soft, soft.ctypes.data, nsoft)
        cap = nsoft // 2 + 1
    else:
        cap = total_bits
    hp_pairs = lib.v224x_host_alloc_pinned(2 * cap) if not on_gpu else None
    flips_buf = np.zeros(4096, np.int64)
    nflips = ctypes.c_int(0)

    def pair_on_host():
        """vdecode.c:101-140 over the whole received stream, once (the flip decisions depend on received symbols only)."""
        return int(lib.v224x_pair_symbols(hp_soft, nsoft, 0, 1 if dontflip else 0, DELAY, hp_pairs, None,
                                          flips_buf.ctypes.data_as(ctypes.c_void_p), flips_buf.size, ctypes.byref(nflips)))

    npairs = pair_on_host() if not on_gpu else total_bits
    flips = [int(x) for x in flips_buf[: nflips.value]]
    segs = v224.segments.plan(npairs, world, W, DELAY)
    me = segs[rank]
    my_lead, my_nout = me.out_first - me.stage_first, me.out_last - me.out_first
    longest = max(s.out_last - s.out_first for s in segs)
    dsyms = torch.empty(2 * (me.out_last - me.stage_first), dtype=torch.uint8, device=dev)
    dbits = torch.empty(longest, dtype=torch.uint8, device=dev)
    if on_gpu:
        tx_range, sym_range = gpu_stream_range(torch, dev, me.stage_first, me.out_last)
        dsyms.copy_(sym_range)
        del sym_range
        hp_pairs_range = lib.v224x_host_alloc_pinned(dsyms.numel())              # the rank's pairs on the host, for the e2e leg
        torch.cuda.synchronize()
        dec.d2h(np.ctypeslib.as_array(ctypes.cast(hp_pairs_range, ctypes.POINTER(ctypes.c_uint8)), (dsyms.numel(),)), dsyms.data_ptr())
    else:
        dec.h2d(dsyms.data_ptr(), np.ctypeslib.as_array(ctypes.cast(hp_pairs, ctypes.POINTER(ctypes.c_uint8)), (2 * cap,))[2 * me.stage_first: 2 * me.out_last])
    torch.cuda.synchronize()
    gathered = [torch.empty(longest, dtype=torch.uint8, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    hp_out = lib.v224x_host_alloc_pinned(npairs) if rank == 0 else None
    out_host = np.ctypeslib.as_array(ctypes.cast(hp_out, ctypes.POINTER(ctypes.c_uint8)), (npairs,)) if rank == 0 else None
    redo_bufs = []

    def load_resident(a, b):
        if (a, b) == (me.stage_first, me.out_last):
            return dsyms, dbits
        # redo of a later range (a failed hand-over): that range's pairs come from the host copy
        s = torch.empty(2 * (b - a), dtype=torch.uint8, device=dev)
        if on_gpu:
            _, sr = gpu_stream_range(torch, dev, a, b)
            s.copy_(sr)
            torch.cuda.synchronize()
        else:
            dec.h2d(s.data_ptr(), np.ctypeslib.as_array(ctypes.cast(hp_pairs, ctypes.POINTER(ctypes.c_uint8)), (2 * cap,))[2 * a: 2 * b])
        o = torch.empty(b - a, dtype=torch.uint8, device=dev)
        redo_bufs.append((s, o))
        return s, o

    reports = []

    def step_device():
        first, bits_t, rep = v224.segments.decode_verified(rd, load_resident, npairs, DELAY, CONV, rank, world, dist, torch, ctrl_device=dev)
        reports.append(rep)
        return rep

    def load_from_host(a, b):
        if (a, b) != (me.stage_first, me.out_last):
            return load_resident(a, b)
        if on_gpu:
            lib.v224x_h2d(dec.h, dsyms.data_ptr(), hp_pairs_range, 2 * (b - a))
        else:
            lib.v224x_h2d(dec.h, dsyms.data_ptr(), ctypes.c_void_p(hp_pairs + 2 * a), 2 * (b - a))
        return dsyms, dbits

    def step_e2e():
        if not on_gpu:
            n2 = pair_on_host()
            assert n2 == npairs
        first, bits_t, rep = v224.segments.decode_verified(rd, load_from_host, npairs, DELAY, CONV, rank, world, dist, torch, ctrl_device=dev)
        if world > 1:
            dist.gather(dbits, gathered, dst=0)
            if rank == 0:
                torch.cuda.synchronize()
                for s, t in zip(segs, gathered):
                    lib.v224x_d2h(dec.h, ctypes.c_void_p(hp_out + s.out_first), t.data_ptr(), s.out_last - s.out_first)
        else:
            lib.v224x_d2h(dec.h, hp_out, dbits.data_ptr(), npairs)
        return rep

    # ---------------- value leg: pairs resident in HBM ----------------
    for _ in range(args.warmup):
        step_device()
    dec.kernel_time_enable(True)
    l0 = dec.stats()["launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dec.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rep_dev = step_device()
    torch.cuda.synchronize()
    ms_dev = dec.timer_stop_ms()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = dec.stats()
    launches = st["launches"] - l0
    acs_ms, acs_launches, acs_passes = dec.kernel_time_ms()
    dec.kernel_time_enable(False)
    seg_rep = rd.reports[-1]

    # ---------------- e2e leg: host buffers through the C ABI ----------------
    step_e2e()
    barrier()
    dec.timer_start()
    for _ in range(args.steps):
        rep_e2e = step_e2e()
    torch.cuda.synchronize()
    ms_e2e = dec.timer_stop_ms()
    barrier()

    # ---------------- checks (untimed) ----------------
    residual = None
    errs = nchk = ntrans = 0
    if rank == 0:
        out = out_host.copy()
        if not on_gpu:
            # the same pair stream on ONE GPU (rank 0), lockstep-verified sequential decode: the N-GPU output must equal it bit for bit
            if world > 1:
                pairs_np = np.ctypeslib.as_array(ctypes.cast(hp_pairs, ctypes.POINTER(ctypes.c_uint8)), (2 * npairs,))
                with v224.Viterbi224(BLOCK + DELAY, device=local_rank) as d1:
                    d1.init(0)
                    one, _ = d1.stream_decode_seg(pairs_np, DELAY, nseg, CONV)
                residual = int((out != one).sum())
            else:
                # N = 1: the plain sequential single-decoder path over the same pairs
                dec.init(0)
                dec.stream_decode_dev(dsyms.data_ptr(), npairs, DELAY, dbits.data_ptr())
                one = np.empty(npairs, np.uint8)
                dec.d2h(one, dbits.data_ptr())
                residual = int((out != one).sum())
            errs, nchk, ntrans = ber_check(out, tx_bits, flips, DELAY, (JUNK - 1) // 2 if args.config == 2 else 0)
    if on_gpu:
        # every rank compares its own range with the transmitted data on the GPU (lag delay + K - 2)
        lag = DELAY + K - 2
        lo = max(0, lag + me.stage_first - me.out_first)                # output i of the range is data bit out_first + i - lag
        off = me.out_first - lag - me.stage_first
        errs = int((dbits[lo:my_nout] != tx_range[off + lo: off + my_nout]).sum().item())
        nchk = int(my_nout - lo)

    t_dev = torch.tensor([ms_dev, ms_e2e, wall * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(launches), float(errs) if (on_gpu or rank == 0) else 0.0, float(nchk) if (on_gpu or rank == 0) else 0.0, float(acs_passes),
                        float(seg_rep["verified"]), float(seg_rep["redone"])], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_dev_max, ms_e2e_max, wall_max = (float(x) for x in t_dev)
    launches_total, errs_total, nchk_total, passes_total, inner_verified, inner_redone = (float(x) for x in tot)

    if rank == 0:
        peak, peak_src, _ = peaks()
        value = npairs * args.steps / (ms_dev_max * 1e-3)
        e2e = npairs * args.steps / (ms_e2e_max * 1e-3)
        persist_launches = max(1, acs_launches // (2 * int(seg_rep["segments"]) + 1))
        achieved = (B_PASS * acs_passes / (acs_ms * 1e-3)) / 1e9 if acs_ms > 0 else None      # GB/s, this rank's kernel
        traffic = None
        tp = os.path.join(ROOT, "profiles", "fused_traffic.json")
        if os.path.exists(tp):
            per_pass = json.load(open(tp)).get("dram_bytes_per_pass")
            traffic = per_pass * acs_passes / persist_launches if per_pass else None      # per launch, like `achieved`
        h2d = int(sum(2 * (s_.out_last - s_.stage_first) for s_ in segs))
        line = {"metric": "decoded_bits_per_s", "value": value, "unit": "bits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True, "scaling": scaling_of(args, world), "vs_baseline": None, "dtype": "u16",
                "data": "synthetic", "config": workload_config(args, world, total_bits),
                "state_updates_per_s": value * (1 << 23),
                "per_gpu": {"value": value / world, "unit": "bits/s", "note": "whole-job value / GPUs: what weak scaling (the same bits per GPU, "
                            "--scaling weak) measures, since the rate does not depend on the stream's length"},
                "e2e": {"value": e2e, "unit": "bits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(npairs),
                        "ms_per_step": ms_e2e_max / args.steps,
                        "includes": "host pairing / phase flip over the whole stream (v224x_pair_symbols), H2D of every rank's pairs from pinned "
                                    "memory, decode, hand-over verification, gather of all decoded bits on rank 0, D2H"},
                "gpu_launches": int(launches_total),
                "roofline": {"bound": "hbm", "kernel": "k_acs_persist", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_pass": B_PASS, "passes_per_launch": acs_passes / persist_launches,
                             "algorithmic_bytes_per_launch": B_PASS * acs_passes / persist_launches,
                             "launches_timed": acs_launches, "passes_timed": acs_passes, "mean_pass_us": 1e3 * acs_ms / max(1, acs_passes),
                             "frac_unfused_equivalent": value / world * B_STAGE_UNFUSED / 1e9 / peak},
                "clocks": clocks,
                "check": {"residual_diffs_vs_one_gpu_decode": residual, "bit_errors_vs_transmitted": int(errs_total), "bits_checked": int(nchk_total),
                          "bits_in_phase_flip_transients": int(ntrans), "phase_flips": flips,
                          "rank_handovers": {k: rep_e2e[k] for k in ("handovers_verified", "ranges_redone", "worst_spread")},
                          "lockstep_handovers_all_ranks": {"verified": int(inner_verified), "redone": int(inner_redone)},
                          "segments_rank0": seg_rep, "wall_ms_per_step": wall_max / args.steps,
                          "passes": {k: st[k] for k in ("fused_passes", "careful_passes", "single_stages", "sat_stages")}}}
        if world == 1 and not args.no_side_rooflines:
            line["roofline_other_kernels"] = side_rooflines(v224, dec, dsyms.data_ptr(), peak)
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sample = 384
            v, kind, dt = run_reference_sample(sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": "bits/s", "cores": threads, "kind": kind,
                                    "sample": f"{threads} decoders (one per host core) x {sample} symbol pairs of the same symdemod-format stream, "
                                              f"update(1)+decodebit(200,0) per bit, {dt:.1f} s wall"}
        print(json.dumps(line), flush=True)

    dec.delete()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
