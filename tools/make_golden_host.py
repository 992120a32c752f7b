#!/usr/bin/env python
"""Generate tests/golden/host/*.npz: what the UNMODIFIED reference host programs print on seeded inputs.

  decode_V_seed21   oracle/_ref/decode_sse -V   (decode.c + viterbi224_sse2.c, reference Makefile:27) on a framed telemetry
                    stream with a junk prefix, 100 symbols lost inside one frame and 5 inserted into another: frame sync search, lock, loss of
                    lock (bad frame), re-acquisition.  ~2.7 s of CPU per frame.
  decode_V_long_seed77  the same program on 60 frames at 3 dB with four disturbances (56 frames found, 5 bad); printout + symbol CRC.
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


def framer_bits(seed=7):
    rng = np.random.default_rng(seed)
    bits = np.concatenate([rng.integers(0, 2, 333, dtype=np.uint8), S.telemetry_bits(6, rng)])
    bits[333 + 3 * 1024 - 5] ^= 1
    return bits


def main():
    os.makedirs(OUT, exist_ok=True)
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
    bits = framer_bits()
    txt = bytes(np.where(bits == 1, ord("1"), ord("0")).astype(np.uint8))
    r = subprocess.run([os.path.join(REF, "framer_ref"), "-r", "512"], input=txt, capture_output=True, env=env, check=True)
    np.savez_compressed(os.path.join(OUT, "framer_seed7.npz"), bits=np.packbits(bits), nbits=bits.size,
                        stdout=np.frombuffer(r.stdout, dtype=np.uint8))
    print(r.stdout.decode()[:200])


if __name__ == "__main__":
    main()
