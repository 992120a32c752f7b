#!/usr/bin/env python
"""Regenerates tests/golden/host/vdecode_long_seed2014.npz: BASELINE config 2 at scale through the UNMODIFIED reference.

The stream: 72 minor frames (73,728 bits) of symdemod-format telemetry at Eb/N0 3 dB (streams.telemetry_stream), an odd
junk prefix of 101 noise symbols (vdecode starts on the wrong symbol phase and flips after the first sync period,
vdecode.c:118-140) and one symbol dropped in mid-stream (a symbol slip: the decoder is fed mis-paired symbols until the
correlator flips the phase again).  `oracle/_ref/vdecode_sse -d 200 -i 8192` (vdecode.c + viterbi224_sse2.c compiled as
they are by oracle/Makefile) decodes it bit by bit -- about 4 CPU-minutes -- and its standard output (one character per
decoded bit, packed here) and standard error (phase-flip notices and the re-encode tally lines) are the fixture.  The
symbols are regenerated from the seed by the tests and checked against the stored CRC.

    python tools/make_golden_vdecode_long.py          (needs /root/reference, i.e. oracle/_ref built here)
"""
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224   # noqa: E402

NBITS = 72 * 1024
EBN0 = 3.0
SEED = 2014
JUNK = 101
SLIP_AT = 70001          # index (in the junk-prefixed stream) of the symbol that is lost in transit
DELAY = 200
INTERVAL = 8192


def stream():
    _, soft = v224.streams.telemetry_stream(NBITS, EBN0, seed=SEED, junk_symbols=JUNK)
    return np.ascontiguousarray(np.delete(soft, SLIP_AT))


def main():
    soft = stream()
    exe = os.path.join(ROOT, "oracle", "_ref", "vdecode_sse")
    env = dict(os.environ, LANG="C")
    r = subprocess.run([exe, "-d", str(DELAY), "-i", str(INTERVAL)], input=soft.tobytes(), capture_output=True, env=env, check=True)
    out = np.frombuffer(r.stdout, dtype=np.uint8)
    assert set(np.unique(out)) <= {ord("0"), ord("1")}
    err = r.stderr.decode().replace(exe, "vdecode")
    dst = os.path.join(ROOT, "tests", "golden", "host", "vdecode_long_seed2014.npz")
    np.savez_compressed(dst, bits_packed=np.packbits(out - ord("0")), nout=np.int64(out.size), stderr=np.array(err),
                        soft_crc=np.uint32(zlib.crc32(soft.tobytes())), nsyms=np.int64(soft.size),
                        params=np.array([NBITS, SEED, JUNK, SLIP_AT, DELAY, INTERVAL], dtype=np.int64), ebn0=np.float64(EBN0))
    print(f"{dst}: {out.size} decoded bits, {err.count('flipping phase')} phase flips, stderr {len(err)} bytes")
    print(err)


if __name__ == "__main__":
    main()
