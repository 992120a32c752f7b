#!/usr/bin/env python
"""bench.py -- decoded bits/s of the B200 viterbi224 decoder on BASELINE.json's config 2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one complete decode of a 1,048,576-bit (1024 minor frames) symdemod-format soft-symbol
stream through the streaming path of vdecode.c (update + decodebit(delay=200, state 0) per bit,
vdecode.c:145-152) in block form: 131072 fused 8-stage ACS passes + batched tracebacks, run as 3 contiguous segments by 3 decoders that one persistent
kernel advances in lockstep (every hand-over between segments verified on the device, output identical to the sequential decode).  The
sync-correlator phase flip (vdecode.c:107-140) is host logic and runs once, before the timed region.

  value : whole-job decoded bits/s with the symbol pairs resident in HBM (device events, max over ranks)
  e2e   : the same through the host-buffer C-ABI call v224x_stream_decode (pinned host memory,
          H2D of the symbols and D2H of the decoded bits inside the timed region)
  roofline : dominant kernel k_acs_persist -- algorithmic bytes per launch (2*16 MiB metrics + 8 MiB
          decisions + 16 symbol bytes) / mean launch duration from CUDA events on the library's stream
  cpu_baseline : the reference's own SSE2 decoder (oracle/_ref, compiled from the unmodified sources) on
          all host cores, one independent stream prefix per core, timed in the same run (rank 0, N=1)

N > 1: one process per GPU (torchrun), the stream is N x 1,048,576 bits cut into time segments with a
warm-up prefix per rank (isee3-decoder_b200/segments.py); no data-path collective; weak scaling.
"""
import argparse
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

NBITS = 1 << 20                 # bits per GPU per step ("1M bits", 1024 minor frames)
DELAY = 200                     # vdecode default decode delay (vdecode.c:44)
BLOCK = 8192                    # stages per update batch; ring = BLOCK + DELAY rows (8.2 GiB)
WARMUP_STAGES = 2048            # leading warm-up of mid-stream segments (N > 1)
SEGMENTS = 3                    # per GPU: decoders advanced in lockstep over contiguous segments of the rank's stream
CONV = 2048                     # stages a late-started decoder gets to converge before its verified hand-over
EBN0_DB = 3.0
SEED = 20141
FK = 8
B_PASS = 2 * (1 << 24) + FK * (1 << 20) + 2 * FK      # algorithmic bytes of one fused pass
B_STAGE_UNFUSED = 34603010                             # bytes per decoded bit of the unfused algorithm


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


def make_workload(rank, world):
    """Rank's share of the stream: pairs (uint8[2*n]) to decode, how many leading outputs are warm-up,
    and the transmitted bits for the BER check."""
    import isee3_decoder_b200 as v224
    S = v224.streams
    rng = np.random.default_rng([SEED, rank])
    bits = S.telemetry_bits(NBITS // S.FRAMEBITS, np.random.default_rng([SEED, 1000 + rank]))
    if rank == 0:
        prev_tail = np.zeros(0, np.uint8)
        state = 0
    else:
        prev = S.telemetry_bits(NBITS // S.FRAMEBITS, np.random.default_rng([SEED, 1000 + rank - 1]))
        prev_tail = prev[-WARMUP_STAGES:]
        hist = prev[-WARMUP_STAGES - 24:-WARMUP_STAGES]
        state = int("".join(map(str, hist)), 2)
    allbits = np.concatenate([prev_tail, bits])
    sym01, _ = S.encode_bits(allbits, state)
    soft = S.awgn_symdemod(sym01, EBN0_DB, rng)
    junk = 0
    if rank == 0:
        # odd junk prefix: vdecode starts on the wrong symbol phase and flips after the first sync period
        _, sigma = S.symdemod_amplitudes(EBN0_DB)
        junk = 101
        soft = np.concatenate([np.clip(128.0 + sigma * rng.standard_normal(junk), 0, 255).astype(np.uint8), soft])
    pairs, flips = v224.vdecode.pair_symbols(soft, return_flips=True)
    return {"pairs": np.ascontiguousarray(pairs.reshape(-1)), "npairs": pairs.shape[0], "skip": prev_tail.size, "bits": allbits,
            "flips": flips, "junk": junk}


def ber_check(out_bits, wl):
    """Decoded output vs transmitted data (lag = delay + K - 2 pairs, vdecode.c:176-177).

    vdecode's sync correlator (vdecode.c:107-140, mirrored exactly on the host) flips the symbol phase whenever the
    out-of-phase correlation peak of a frame beats the in-phase one.  That happens by design behind rank 0's odd junk
    prefix, and -- at 3 dB, as in the reference -- now and then as a false alarm that the next frame corrects.  Every
    flip drops one symbol: between a flip and its correction the decoder is fed mis-paired symbols (garbage out, the
    reference prints the same), and after the correction the output is one more pair behind the transmitted data.
    So the comparison is made window by window over a small set of alignments; windows that match no alignment are the
    flip transients and are reported separately, not counted as decoder errors.
    Returns (bit errors, bits compared, bits inside flip transients)."""
    n = wl["npairs"]
    lag = DELAY + 22
    start = 8192          # well past the initial flip / warm-up transient
    win = 4096
    base = (wl["junk"] + 1) // 2
    nf = len(wl["flips"])
    # every corrected false alarm drops one pair (the output runs one pair AHEAD of where it was); the junk prefix delays it
    shifts = list(range(-(nf + 1), base + 2))
    errs = compared = transient = 0
    for a in range(start, n, win):
        idx = np.arange(a, min(n, a + win))
        best = None
        for shift in shifts:
            src = idx - lag - shift
            ok = (src >= 0) & (src < wl["bits"].size)
            if ok.sum() < idx.size // 2:
                continue
            e = int((out_bits[idx[ok]] != wl["bits"][src[ok]]).sum())
            if best is None or e < best[0]:
                best = (e, int(ok.sum()))
        if best is None:
            continue
        if best[0] > best[1] // 128:          # no alignment fits cleanly: wrong symbol phase / re-acquisition after a flip
            transient += best[1]
        else:
            errs += best[0]
            compared += best[1]
    return errs, compared, transient


def run_reference_sample(sample_bits, threads):
    """The reference's SSE2 decoder on `threads` host cores, one independent decoder per core, each running
    the vdecode.c:145-152 loop (update(1) + decodebit(200, 0)) over `sample_bits` pairs.  Returns (bits/s, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    import isee3_decoder_b200 as v224
    Dec = pyoracle.best_cpu_decoder()
    kind = "reference" if Dec is pyoracle.RefSSE2 else "port"
    _, soft = v224.streams.telemetry_stream(sample_bits, EBN0_DB, seed=SEED + 7)
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


def persistent_launches(launches, nseg):
    """Each persistent launch over nseg decoders is preceded by two small bookkeeping kernels per decoder
    (k_build_passtab, k_persist_begin) that are inside the timed ACS region but move ~1 KiB per pass."""
    return max(1, launches // (2 * nseg + 1))


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    sample = 384
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, dt = run_reference_sample(sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    ms = 1e3 * statistics.mean(dt for _, dt in vals)
    sample_desc = (f"{threads} decoders (one per host core) x {sample} symbol pairs of the same symdemod-format stream, "
                   f"update(1)+decodebit(200,0) per bit; len {DELAY + 1} ring")
    line = {"impl": "reference", "metric": "decoded_bits_per_s", "value": value, "unit": "bits/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "bits/s", "cores": threads, "kind": kind, "sample": sample_desc},
            "e2e": {"value": value, "unit": "bits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "state_updates_per_s": value * (1 << 23), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "BASELINE config 2: vdecode streaming decode of a synthetic symdemod-format telemetry stream, "
                        f"{NBITS} bits (1024 minor frames) per GPU, Eb/N0 {EBN0_DB} dB, decode delay {DELAY}, automatic symbol-phase flip "
                        "(odd junk prefix on rank 0)",
            "bits_per_gpu": NBITS, "decode_delay": DELAY, "stages_per_pass": FK, "block_stages": BLOCK,
            "parallelism": (f"time-segmented x{n} GPUs, warm-up {WARMUP_STAGES} stages; " if n > 1 else "single GPU; ")
                           + f"per GPU {SEGMENTS} decoders in lockstep over contiguous segments, hand-overs verified on the device "
                             f"(warm-up {DELAY}+{CONV} stages each)",
            "cache": "the decision rings (8.2 GiB per decoder) are written once per stage and are far larger than the 126 MB L2; the "
                     "16 MiB path-metric buffers (3 per decoder) are re-read by the next pass by construction (no artificial L2 "
                     "flush possible without changing the algorithm)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--segments", type=int, default=SEGMENTS, help="decoders advanced in lockstep per GPU (1 = sequential)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import isee3_decoder_b200 as v224
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL announces its version on stdout when the first communicator is built; stdout carries exactly one JSON
        # line, so the file descriptor points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
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

    wl = make_workload(rank, world)
    n = wl["npairs"]
    dec = v224.Viterbi224(BLOCK + DELAY, device=local_rank)
    lib = dec.lib
    dsyms = dec.dev_alloc(2 * n)
    dbits = dec.dev_alloc(n)
    dec.h2d(dsyms, wl["pairs"])
    # pinned host buffers for the e2e leg
    import ctypes
    hp_syms = lib.v224x_host_alloc_pinned(2 * n)
    hp_bits = lib.v224x_host_alloc_pinned(n)
    ctypes.memmove(hp_syms, wl["pairs"].ctypes.data, 2 * n)

    nseg = max(1, args.segments)
    seg_rep = {}

    def step_device():
        dec.init(0) if rank == 0 else dec.init_uniform(5000, -1)
        seg_rep.update(dec.stream_decode_seg_dev(dsyms, n, DELAY, dbits, nseg, CONV))

    def step_e2e():
        dec.init(0) if rank == 0 else dec.init_uniform(5000, -1)
        rep = v224.binding.SegReport()
        r = lib.v224x_stream_decode_seg(dec.h, hp_syms, n, DELAY, hp_bits, nseg, CONV, ctypes.byref(rep))
        assert r >= 0, lib.v224x_last_error()
        return r

    # ---------------- reference output of this rank: the sequential block decode (untimed) ----------------
    dec.init(0) if rank == 0 else dec.init_uniform(5000, -1)
    dec.stream_decode_dev(dsyms, n, DELAY, dbits)
    out_seq = np.empty(n, np.uint8)
    dec.d2h(out_seq, dbits)

    # ---------------- value leg: inputs resident in HBM ----------------
    for _ in range(args.warmup):
        step_device()
    out = np.empty(n, np.uint8)
    dec.d2h(out, dbits)
    seg_same = bool(np.array_equal(out, out_seq))
    errs, nchk, ntrans = ber_check(out, wl)
    dec.kernel_time_enable(True)
    l0 = dec.stats()["launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dec.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    ms_dev = dec.timer_stop_ms()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = dec.stats()
    launches = st["launches"] - l0
    acs_ms, acs_launches, acs_passes = dec.kernel_time_ms()
    dec.kernel_time_enable(False)

    # ---------------- e2e leg: host buffers through the C ABI ----------------
    step_e2e()
    barrier()
    dec.timer_start()
    for _ in range(args.steps):
        step_e2e()
    ms_e2e = dec.timer_stop_ms()
    barrier()
    out2 = np.ctypeslib.as_array(ctypes.cast(hp_bits, ctypes.POINTER(ctypes.c_uint8)), (n,)).copy()
    same = bool(np.array_equal(out, out2))

    t_dev = torch.tensor([ms_dev, ms_e2e, wall * 1e3], dtype=torch.float64, device="cuda" if world > 1 else "cpu")
    tot = torch.tensor([float(n - wl["skip"]), float(launches), float(errs), float(nchk), float(ntrans), float(len(wl["flips"]))],
                       dtype=torch.float64, device=t_dev.device)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_dev_max, ms_e2e_max, wall_max = (float(x) for x in t_dev)
    bits_total, launches_total, errs_total, nchk_total, ntrans_total, nflips_total = (float(x) for x in tot)

    if rank == 0:
        peak, peak_src, _ = peaks()
        nseg_used = int(seg_rep.get("segments", 1))
        value = bits_total * args.steps / (ms_dev_max * 1e-3)
        e2e = bits_total * args.steps / (ms_e2e_max * 1e-3)
        achieved = (B_PASS * acs_passes / (acs_ms * 1e-3)) / 1e9 if acs_ms > 0 else None      # GB/s, this rank's kernel
        traffic = None
        tp = os.path.join(ROOT, "profiles", "fused_traffic.json")
        if os.path.exists(tp):
            per_pass = json.load(open(tp)).get("dram_bytes_per_pass")
            traffic = per_pass * acs_passes / persistent_launches(acs_launches, nseg_used) if per_pass else None      # per launch, like `achieved`
        line = {"metric": "decoded_bits_per_s", "value": value, "unit": "bits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16",
                "data": "synthetic", "config": workload_config(world),
                "state_updates_per_s": value * (1 << 23),
                "e2e": {"value": e2e, "unit": "bits/s", "h2d_bytes_per_step": int(2 * n), "d2h_bytes_per_step": int(n),
                        "ms_per_step": ms_e2e_max / args.steps, "output_identical_to_device_leg": same},
                "gpu_launches": int(launches_total),
                "roofline": {"bound": "hbm", "kernel": "k_acs_persist", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_pass": B_PASS, "passes_per_launch": acs_passes / persistent_launches(acs_launches, nseg_used),
                             "algorithmic_bytes_per_launch": B_PASS * acs_passes / persistent_launches(acs_launches, nseg_used),
                             "launches_timed": acs_launches, "passes_timed": acs_passes, "mean_pass_us": 1e3 * acs_ms / max(1, acs_passes),
                             "frac_unfused_equivalent": value / world * B_STAGE_UNFUSED / 1e9 / peak},
                "clocks": clocks,
                "check": {"bit_errors_vs_transmitted": int(errs_total), "bits_checked": int(nchk_total),
                          "bits_in_phase_flip_transients": int(ntrans_total), "phase_flips_all_ranks": int(nflips_total),
                          "phase_flips_rank0": wl["flips"],
                          "segmented_output_identical_to_sequential_rank0": seg_same, "segments_rank0": seg_rep,
                          "wall_ms_per_step": wall_max / args.steps,
                          "passes": {k: st[k] for k in ("fused_passes", "careful_passes", "single_stages", "sat_stages")}}}
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sample = 384
            v, kind, dt = run_reference_sample(sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": "bits/s", "cores": threads, "kind": kind,
                                    "sample": f"{threads} decoders (one per host core) x {sample} symbol pairs of the same symdemod-format stream, "
                                              f"update(1)+decodebit(200,0) per bit, {dt:.1f} s wall"}
        print(json.dumps(line), flush=True)

    lib.v224x_host_free_pinned(hp_syms)
    lib.v224x_host_free_pinned(hp_bits)
    dec.dev_free(dsyms)
    dec.dev_free(dbits)
    dec.delete()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
