#!/usr/bin/env python
"""Generate tests/golden/host/*.npz: what the UNMODIFIED reference host programs print on seeded inputs.

  decode_V_seed21   oracle/_ref/decode_sse -V   (decode.c + viterbi224_sse2.c, reference Makefile:27) on a framed telemetry
                    stream with a junk prefix, 100 symbols lost inside one frame and 5 inserted into another: frame sync search, lock, loss of
                    lock (bad frame), re-acquisition.  ~2.7 s of CPU per frame.
  decode_V_long_seed77  the same program on 60 frames at 3 dB with four disturbances (56 frames found, 5 bad); printout + symbol CRC.
  decode_hybrid_*   decode_sse in its default mode (Fano first, Viterbi fallback) and with -p -n; printout + symbol CRC.
  fano_cases        fano() / gen_met() / format_hms() of the reference library (oracle/_ref/libv224_reffano.so) on seeded frames.
  framer_seed7      oracle/_ref/framer_ref -r 512 (framer.c, reference Makefile:40) on a decoded-bit stream with a junk
                    prefix and one corrupted sync word.

Run in the build container (the reference checkout does not exist on the GPU box):   python tools/make_golden_host.py
The fixtures pin isee3-decoder_b200/bin/decode_block (GPU tier) and vdecode_block -f (CPU tier: -b mode)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224      # noqa: E402  (stream generators only)

S = v224.streams
OUT = os.path.join(ROOT, "tests", "golden", "host")
REF = os.path.join(ROOT, "oracle", "_ref")


def decode_stream(seed=21, nframes=9, ebn0=4.0, junk=777, cut_frame=3, cut_at=900, cut=100, ins_frame=6, ins_at=1500, ins=5):
    """Soft symbols: junk, then `nframes` minor frames; `cut` symbols are removed from frame `cut_frame` (0-based) and
    `ins` noise symbols inserted into frame `ins_frame`: two losses of frame lock with re-acquisition.
    (The reference overruns its 4096-symbol buffer when the re-acquired sync position lies beyond 2014, decode.c:48,183-188
    -- e.g. after a slip of fewer than 34 symbols it dies of a smashed stack; the fixtures stay clear of that.)"""
    _, soft = S.telemetry_stream(nframes * S.FRAMEBITS, ebn0, seed=seed, junk_symbols=junk)
    a = junk + cut_frame * 2 * S.FRAMEBITS + cut_at
    b = junk + ins_frame * 2 * S.FRAMEBITS + ins_at
    extra = np.random.default_rng(seed + 1).integers(0, 256, ins, dtype=np.uint8)
    return np.concatenate([soft[:a], soft[a + cut:b], extra, soft[b:]])


def decode_stream_long():
    """60 frames at 3 dB, junk prefix 1234, symbols cut (40, 1000) and inserted (7, 300) inside four frames, truncated end."""
    _, soft = S.telemetry_stream(60 * 1024, 3.0, seed=77, junk_symbols=1234)
    rng = np.random.default_rng(78)
    for frame, at, cut, ins in ((50, 100, 1000, 0), (37, 1999, 0, 300), (21, 5, 40, 0), (9, 1024, 0, 7)):      # back to front
        a = 1234 + frame * 2048 + at
        soft = np.concatenate([soft[:a], rng.integers(0, 256, ins, dtype=np.uint8), soft[a + cut:]])
    return soft[:-700]


def hybrid_cases():
    """(fixture name, decode flags, symbols) for the Fano-first mode: the disturbed 7-frame stream (a Fano failure right
    after a loss of lock is NOT retried with Viterbi), 40 frames at 1.75 dB (several Viterbi fallbacks) and 40 frames at
    1.5 dB with -p (Viterbi also without lock)."""
    yield "decode_hybrid_seed21", [], decode_stream()
    yield "decode_hybrid_1p75dB_seed5", [], S.telemetry_stream(40 * 1024, 1.75, seed=5, junk_symbols=2500)[1]
    yield "decode_hybrid_p_1p5dB_seed7", ["-p", "-n"], S.telemetry_stream(40 * 1024, 1.5, seed=7, junk_symbols=2500)[1]


def framer_bits(seed=7):
    rng = np.random.default_rng(seed)
    bits = np.concatenate([rng.integers(0, 2, 333, dtype=np.uint8), S.telemetry_bits(6, rng)])
    bits[333 + 3 * 1024 - 5] ^= 1
    return bits


FANO_TABLES = [(81.64965809277261, 57.73502691896258, 0.5, 8.0), (24.0, 16.99, 0.5, 8.0), (100.0, 20.0, 0.5, 16.0), (30.0, 60.0, 0.0, 4.0),
               (120.0, 3.0, 0.5, 8.0)]        # (signal, noise, bias, scale); the first is decode.c's own (decode.c:121-137)


def fano_cases():
    """(name, symbols, nbits, table index, delta, maxcycles, start state, tail bits): frames at several Eb/N0 around the
    sequential decoder's threshold, so that clean decodes, long searches and timeouts all occur."""
    sync24 = S.SYNCWORD & 0xFFFFFF
    for i, ebn0 in enumerate([6.0, 3.0, 2.5, 2.0, 2.0, 1.5, 1.5, 1.0, 0.0]):
        rng = np.random.default_rng(900 + i)
        bits = S.telemetry_bits(1, rng)
        sym01, _ = S.encode_bits(bits, sync24)
        yield (f"frame_{ebn0}dB_{i}", S.awgn_symdemod(sym01, ebn0, rng), 1024, 0, 32, 100, sync24, sync24)
    rng = np.random.default_rng(950)
    data = np.zeros(40, np.uint8); data[:37] = rng.integers(0, 256, 37, dtype=np.uint8)
    sym01, _ = S.encode(data, 0)
    yield ("vtest_320_3dB", S.awgn_vtest(sym01, 3.0, rng), 320, 1, 16, 1000, 0, 0)
    yield ("vtest_320_timeout", S.awgn_vtest(sym01, -1.0, rng), 320, 1, 16, 3, 0, 0)


def ref_fano_lib():
    import ctypes
    lib = ctypes.CDLL(os.path.join(REF, "libv224_reffano.so"))
    lib.fano.restype = ctypes.c_int
    lib.fano.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint, ctypes.c_void_p, ctypes.c_int, ctypes.c_ulong, ctypes.c_ulonglong, ctypes.c_ulonglong]
    lib.gen_met.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
    lib.format_hms.restype = ctypes.c_char_p
    lib.format_hms.argtypes = [ctypes.c_double]
    return lib


def run_fano(fn, table, syms, nbits, delta, maxcycles, start, tail):
    """fn = fano() of the reference library or shim_fano of tests/emu/host_shim.cpp (same argument list)."""
    import ctypes
    metric, cycles = ctypes.c_ulong(0), ctypes.c_ulong(0)
    data = np.zeros(nbits // 8, np.uint8)
    syms = np.ascontiguousarray(syms)
    r = fn(ctypes.byref(metric), ctypes.byref(cycles), data.ctypes.data, syms.ctypes.data, nbits, table.ctypes.data, delta, maxcycles, start, tail)
    return int(r), int(metric.value), int(cycles.value), data


def main():
    os.makedirs(OUT, exist_ok=True)
    # the sequential decoder, its metric tables and the time format, straight from the reference library
    lib = ref_fano_lib()
    tables = []
    for sig, noise, bias, scale in FANO_TABLES:
        t = np.zeros((2, 256), np.int32)
        lib.gen_met(t.ctypes.data, sig, noise, bias, scale)
        tables.append(t)
    rows, datas = [], []
    for name, syms, nbits, ti, delta, maxc, start, tail in fano_cases():
        r, metric, cycles, data = run_fano(lib.fano, tables[ti], syms, nbits, delta, maxc, start, tail)
        rows.append((r, metric, cycles))
        datas.append(np.pad(data, (0, 128 - data.size)))
        print(f"fano {name:22s} bits {r:5d} metric {metric:8d} cycles {cycles:8d}")
    times = [0.0, 9.9996, 59.9994, 61.5, 3599.9999, 3600.0, 86399.5, 86400.0, 123456.789, 1e6 / 3, 8969 / 1024.0]
    np.savez_compressed(os.path.join(OUT, "fano_cases.npz"), tables=np.array(tables), results=np.array(rows, dtype=np.uint64), data=np.array(datas),
                        times=np.array(times), hms=np.array([lib.format_hms(t).decode() for t in times]))
    soft = decode_stream()
    env = dict(os.environ, LANG="C")
    r = subprocess.run([os.path.join(REF, "decode_sse"), "-V"], input=soft.tobytes(), capture_output=True, env=env, check=True)
    np.savez_compressed(os.path.join(OUT, "decode_V_seed21.npz"), symbols=soft, stdout=np.frombuffer(r.stdout, dtype=np.uint8),
                        argv0=np.frombuffer(os.path.join(REF, "decode_sse").encode(), dtype=np.uint8))
    print(r.stdout.decode()[:400])
    # the long stream: symbols are regenerated from the seeds by the test (CRC stored), only the printout is kept (~2.5 min of CPU)
    import zlib
    soft = decode_stream_long()
    r = subprocess.run([os.path.join(REF, "decode_sse"), "-V"], input=soft.tobytes(), capture_output=True, env=env, check=True)
    np.savez_compressed(os.path.join(OUT, "decode_V_long_seed77.npz"), symbols_crc=np.uint32(zlib.crc32(soft.tobytes())),
                        stdout=np.frombuffer(r.stdout, dtype=np.uint8))
    print(r.stdout.count(b"Frame "), "frames,", r.stdout.count(b"(bad)"), "bad")
    # Fano first, Viterbi fallback (the reference's default mode): only the frames Fano gives up on cost CPU minutes
    for name, flags, soft in hybrid_cases():
        r = subprocess.run([os.path.join(REF, "decode_sse")] + flags, input=soft.tobytes(), capture_output=True, env=env, check=True)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), symbols_crc=np.uint32(zlib.crc32(soft.tobytes())), flags=np.array(flags, dtype="U8"),
                            stdout=np.frombuffer(r.stdout, dtype=np.uint8))
        print(name, flags, r.stdout.count(b"Frame "), "frames,", r.stdout.count(b"with Viterbi"), "by Viterbi,", r.stdout.count(b"(bad)"), "bad")
    bits = framer_bits()
    txt = bytes(np.where(bits == 1, ord("1"), ord("0")).astype(np.uint8))
    r = subprocess.run([os.path.join(REF, "framer_ref"), "-r", "512"], input=txt, capture_output=True, env=env, check=True)
    np.savez_compressed(os.path.join(OUT, "framer_seed7.npz"), bits=np.packbits(bits), nbits=bits.size,
                        stdout=np.frombuffer(r.stdout, dtype=np.uint8))
    print(r.stdout.decode()[:200])


if __name__ == "__main__":
    main()
