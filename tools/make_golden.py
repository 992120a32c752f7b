#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference decoder (viterbi224_sse2.c,
compiled from /root/reference into oracle/_ref by oracle/Makefile) on seeded inputs.

Run in the build container (the reference checkout does not exist on the GPU box):
    python tools/make_golden.py
The fixtures are small (symbols + CRCs), committed, and are what pins both the CPU oracle
(tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import isee3_decoder_b200 as v224      # noqa: E402  (stream generators only; no GPU needed)
import pyoracle                        # noqa: E402
from scripts import run_script, save_case   # noqa: E402

S = v224.streams
GOLD = os.path.join(ROOT, "tests", "golden")


def cases():
    # 1. vtest224 BER frame, 3 dB (config 1 in miniature): init/update/chainback  (vtest224.c:116-118)
    data, syms = S.vtest_frame(256, 3.0, seed=1)
    yield ("awgn3db_256", [["create", 256], ["init", 0], ["update", 0, 256], ["chainback", 256, 0], ["minmax"]], syms,
           {"data_hex": data.tobytes().hex(), "ebn0": 3.0, "style": "vtest"})

    # 2. 1 dB, long enough for renormalisations, fed in ragged chunks (return value = renorm count per call)
    data, syms = S.vtest_frame(648, 1.0, seed=2)
    yield ("awgn1db_648_chunks", [["create", 648], ["init", 0], ["update", 0, 100], ["update", 100, 1], ["update", 101, 7],
                                  ["update", 108, 8], ["minmax"], ["update", 116, 200], ["update", 316, 332], ["minmax"],
                                  ["chainback", 648, 0]], syms, {"data_hex": data.tobytes().hex(), "ebn0": 1.0, "style": "vtest"})

    # 3. vtest224 time-trial input: all-erasure symbols (vtest224.c:165); first renormalisation at stage 227
    syms = np.full(2 * 304, 128, dtype=np.uint8)
    yield ("erasure_304", [["create", 304], ["init", 0], ["update", 0, 304], ["minmax"], ["chainback", 304, 0]], syms, {"style": "erasure"})

    # 4. decode.c:220-222 pattern: known start and end state 0x819fbe (masked to 23 bits inside)
    rng = np.random.default_rng(4)
    bits = rng.integers(0, 2, 128, dtype=np.uint8)
    bits[-24:] = [(0x819FBE >> (23 - i)) & 1 for i in range(24)]
    sym01, st = S.encode_bits(bits, 0x819FBE)
    syms = S.awgn_symdemod(sym01, 4.0, rng)
    yield ("frame_syncstate_128", [["create", 128], ["init", 0x819FBE], ["update", 0, 128], ["chainback", 128, 0x819FBE], ["minmax"]],
           syms, {"bits_hex": np.packbits(bits).tobytes().hex(), "endstate": int(st), "style": "symdemod"})

    # 5. streaming: per-bit update + decodebit (vdecode.c:145-152), then best-path and word variants
    bits, syms = S.telemetry_stream(320, 4.0, seed=5)
    script = [["create", 65], ["init", 0]]
    for i in range(320):
        script += [["update", i, 1], ["decodebit", 64, 0]]
    script += [["decodebit", 64, -1], ["decodeword", 64, 0], ["decodeword", 40, -1], ["decodebit", 0, 0], ["minmax"]]
    yield ("stream_d64_320", script, syms, {"bits_hex": np.packbits(bits).tobytes().hex(), "style": "symdemod"})

    # 6. ring wrap: len 40, 200 stages in odd chunks, traceback across the wrap, chainback with nbits > len
    data, syms = S.vtest_frame(200, 2.0, seed=6)
    script = [["create", 40], ["init", 0]]
    pos = 0
    for n in [13, 8, 8, 3, 16, 24, 9, 40, 41, 38]:
        script += [["update", pos, n], ["decodebit", 32, 0], ["decodeword", 39, 0]]
        pos += n
    script += [["chainback", 200, 0], ["chainback", 37, 5], ["minmax"]]
    yield ("ringwrap_len40_200", script, syms, {"style": "vtest"})

    # 7. strong symbols, all-zero data: state 0 is the best path and the metric spread is large, so the
    #    reference's int16 adds saturate before the renormalisation trigger fires
    rng = np.random.default_rng(7)
    sym01, _ = S.encode_bits(np.zeros(480, np.uint8), 0)
    syms = S.awgn_symdemod(sym01, 7.0, rng)
    yield ("strong_zero_480", [["create", 480], ["init", 0], ["update", 0, 240], ["minmax"], ["update", 240, 240], ["minmax"],
                               ["chainback", 480, 0]], syms, {"style": "symdemod", "ebn0": 7.0})

    # 8. two frames through one handle (re-init resets ring position and renormals; hybridtest.c:186-193)
    d1, s1 = S.vtest_frame(96, 2.0, seed=8)
    d2, s2 = S.vtest_frame(96, 2.0, seed=9)
    syms = np.concatenate([s1, s2])
    yield ("two_frames_96", [["create", 96], ["init", 0], ["update", 0, 96], ["chainback", 96, 0], ["init", 0], ["update", 96, 96],
                             ["chainback", 96, 0], ["minmax"]], syms, {"data_hex": (d1.tobytes() + d2.tobytes()).hex(), "style": "vtest"})


    # 9. forced int16 saturation: a synthetic mid-stream state whose metrics sit just under SHRT_MAX while
    #    state 0 (the renormalisation trigger) is still low -- _mm_adds_epi16 clips for many stages
    data, syms = S.vtest_frame(72, 2.0, seed=10)
    yield ("saturation_forced_72", [["create", 80], ["init", 0], ["set_state", 10, 27000, 32700, [[0, 15000], [1 << 22, 15000]], 123456, 3],
                                    ["update", 0, 1], ["minmax"], ["update", 1, 7], ["update", 8, 8], ["update", 16, 56], ["minmax"],
                                    ["decodebit", 60, -1], ["chainback", 75, 0]], syms, {"style": "vtest", "note": "rows 3..74 written"})


def main():
    if not pyoracle.have_ref():
        pyoracle.build(ref=True)
    assert pyoracle.have_ref(), "the reference could not be compiled (is /root/reference present?)"
    os.makedirs(GOLD, exist_ok=True)
    for name, script, syms, meta in cases():
        out = run_script(lambda n: pyoracle.RefSSE2(n), script, syms)
        meta = dict(meta, generator="tools/make_golden.py", source="viterbi224_sse2.c (unmodified, oracle/_ref/libv224_sse2.so)")
        save_case(os.path.join(GOLD, name + ".npz"), name, script, syms, out, meta)
        nren = sum(r[1] for r in out["results"] if r[0] == "update")
        print(f"{name:28s} ops {len(script):4d} syms {syms.size:5d} renorms {nren} metrics [{out['final']['metrics_min']}, {out['final']['metrics_max']}]")


if __name__ == "__main__":
    main()

# tests/golden/vdecode_flip_seed11.npy is the stdout of the reference's own vdecode binary
# (oracle/_ref/vdecode_sse -d 64 -q) on streams.telemetry_stream(3*1024, 6.0, seed=11, junk_symbols=101):
#   python - <<'PY'
#   import subprocess, numpy as np, isee3_decoder_b200 as v, sys; sys.path.insert(0, 'oracle'); import pyoracle
#   _, soft = v.streams.telemetry_stream(3*1024, 6.0, seed=11, junk_symbols=101)
#   out = subprocess.run([pyoracle.REF_VDECODE, '-d', '64', '-q'], input=soft.tobytes(), capture_output=True).stdout
#   np.save('tests/golden/vdecode_flip_seed11.npy', np.frombuffer(out, dtype=np.uint8))
#   PY
