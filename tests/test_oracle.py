"""CPU tier: pin the oracle (oracle/v224_oracle.c) against
  (a) the golden fixtures recorded from the unmodified reference (tools/make_golden.py),
  (b) the unmodified reference itself when oracle/_ref was built here,
  (c) the reference's one hard-coded golden constant, sync_vector[34] (vdecode.c:27-30),
and check the host emulation of the fused CUDA pass against the oracle."""
import ctypes
import os

import numpy as np
import pytest

import isee3_decoder_b200 as v224
import pyoracle
from scripts import run_script, compare_outcomes

S = v224.streams

# vdecode.c:27-30 (also decode.c:37-40, framer.c:30-33): last 34 encoded symbols of the sync word
SYNC_VECTOR = [0, 1, 1, 1, 1, 1, 1, 0, 1, 0, 1, 1, 1, 1, 0, 0, 1,
               1, 0, 0, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]


def test_sync_vector_pins_polynomials_and_bit_order(built):
    assert S.sync_vector().tolist() == SYNC_VECTOR
    # same through the oracle's C encoder (encode.c:17-35 restatement)
    L = pyoracle.Oracle.lib()
    data = np.array([(S.SYNCWORD >> (8 * i)) & 0xFF for i in range(4, -1, -1)], dtype=np.uint8)
    out = np.zeros(80, dtype=np.uint8)
    st = L.oracle_encode(out.ctypes.data_as(ctypes.c_void_p), data.ctypes.data_as(ctypes.c_void_p), 5, 0)
    assert out[-34:].tolist() == SYNC_VECTOR
    assert st == (S.SYNCWORD & 0xFFFFFF)


def test_numpy_encoder_equals_oracle_encoder(built):
    L = pyoracle.Oracle.lib()
    rng = np.random.default_rng(0)
    for start in (0, 0xABCDEF, 0x819FBE):
        data = rng.integers(0, 256, 64, dtype=np.uint8)
        out = np.zeros(16 * 64, dtype=np.uint8)
        st = L.oracle_encode(out.ctypes.data_as(ctypes.c_void_p), data.ctypes.data_as(ctypes.c_void_p), 64, start)
        sym, st2 = S.encode(data, start)
        assert np.array_equal(sym, out)
        assert st == st2


GOLDEN = sorted(f[:-4] for f in os.listdir(os.path.join(os.path.dirname(__file__), "golden")) if f.endswith(".npz"))


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_reference_golden(built, golden_cases, name):
    case = golden_cases[name]
    got = run_script(lambda n: pyoracle.Oracle(n), case["script"], case["syms"])
    compare_outcomes(got, case["outcome"], f"oracle vs golden {name}")


def test_known_answer_decoded_equals_transmitted(golden_cases):
    """vtest224.c:123-130: at 3 dB the frame decodes to the transmitted data."""
    case = golden_cases["awgn3db_256"]
    decoded = [r for r in case["outcome"]["results"] if r[0] == "chainback"][0][1]
    assert decoded == case["meta"]["data_hex"]
    case = golden_cases["stream_d64_320"]
    bits = np.unpackbits(np.frombuffer(bytes.fromhex(case["meta"]["bits_hex"]), dtype=np.uint8))
    outs = np.array([r[1] for r in case["outcome"]["results"] if r[0] == "decodebit"][:320])
    # output i is data bit i - 64 - 22 (delay + K - 2 lag, vdecode.c:176-177)
    lag = 64 + 22
    assert np.array_equal(outs[lag:], bits[:320 - lag])


def _need_ref(path=None):
    """Skip at RUN time (after the `built` fixture had its chance to compile the reference), not at collection time."""
    if not (os.path.exists(path) if path else pyoracle.have_ref()):
        pytest.skip("reference objects not built here (no reference checkout at build time)")


def test_oracle_equals_reference_on_fresh_input(built):
    _need_ref()
    rng = np.random.default_rng(99)
    n = 72
    syms = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    script = [["create", 50], ["init", 77], ["update", 0, 30], ["decodebit", 20, -1], ["update", 30, 42], ["chainback", 72, 3], ["minmax"]]
    a = run_script(lambda k: pyoracle.Oracle(k), script, syms)
    b = run_script(lambda k: pyoracle.RefSSE2(k), script, syms)
    compare_outcomes(a, b, "oracle vs reference")


def test_sse2_and_portable_reference_agree(built):
    """the two reference builds decode identically despite tie-break/bias differences."""
    _need_ref()
    data, syms = S.vtest_frame(96, 1.0, seed=21)
    with pyoracle.RefSSE2(96) as a, pyoracle.RefPort(96) as b:
        a.init(0); b.init(0)
        a.update_blk(syms, 96); b.update_blk(syms, 96)
        assert np.array_equal(a.chainback(96, 0), b.chainback(96, 0))


def test_channel_restatement_equals_reference_simulate(built):
    """sim.c:17-51: same srandom seed -> same bytes; and the bytes follow the CDF-bin rule used by streams.awgn_vtest."""
    _need_ref(pyoracle.REF_UTIL_SO)
    L = pyoracle.Oracle.lib()
    R = ctypes.CDLL(pyoracle.REF_UTIL_SO)
    R.setup_channel.argtypes = [ctypes.c_double, ctypes.c_double]
    R.simulate.restype = ctypes.c_ubyte
    sigma = float(S.vtest_noise_sigma(3.0))
    assert abs(sigma - 16.99) < 0.01
    R.setup_channel(24.0, sigma)
    L.oracle_setup_channel(24.0, sigma)
    bits = np.random.default_rng(5).integers(0, 2, 4000)
    R.srandom(42)
    a = [R.simulate(int(b)) for b in bits]
    L.oracle_srandom(42)
    b = [L.oracle_simulate(int(x)) for x in bits]
    assert a == b
    mean1 = np.mean([x for x, d in zip(a, bits) if d == 1])
    mean0 = np.mean([x for x, d in zip(a, bits) if d == 0])
    assert abs(mean1 - 152) < 1.5 and abs(mean0 - 104) < 1.5


def _emu(name="libfused_emu.so"):
    so = os.path.join(os.path.dirname(__file__), "emu", "_build", name)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("lib,seed,warm", [("libfused_emu.so", 7, 40), ("libfused_emu.so", 8, 3), ("libfused_emu.so", 9, 25),
                                           ("libfused_emu_t32.so", 7, 40), ("libfused_emu_t32.so", 10, 5),
                                           ("libfused_emu_t32q1.so", 7, 40), ("libfused_emu_t32q1.so", 11, 9),
                                           ("libfused_emu_addforms.so", 7, 40), ("libfused_emu_addforms.so", 12, 2)])
def test_fused_pass_index_algebra_against_oracle(built, lib, seed, warm):
    """Host emulation of the fused kernel's arithmetic core (same header, same per-thread data movement and thread
    maps of both register rounds), for both tile widths the library builds (64 columns: lockstep decoders, 32 columns: a
    decoder running alone): metrics, all eight decision rows (through fused_bit_address) and state-0 tracking equal the
    oracle.  "addforms": every pair with the three-input form of the decision words and the packed-halves form of c + x
    (the library uses them for the pairs that balance the two integer pipes; a half that overflowed would show here)."""
    e = _emu(lib)
    assert e.emu_tile_cols() == (32 if "t32" in lib else 64) and e.emu_nq() == (1 if "q1" in lib else 2)
    fmt_base = e.emu_rowfmt_base()
    rng = np.random.default_rng(seed)
    syms = rng.integers(40, 216, 2 * (warm + 8), dtype=np.uint8)
    with pyoracle.Oracle(warm + 8) as o:
        o.init(0)
        o.update_blk(syms, warm)
        m0 = o.get_metrics()
        P = (m0.astype(np.int32) + 32768).astype(np.uint16)
        sub = int(P.min())
        newP = np.zeros(1 << 23, np.uint16)
        rows = np.zeros((8, 1 << 18), np.uint32)
        stats = np.zeros(19, np.uint32)
        vp = ctypes.c_void_p
        e.emu_fused_pass(P.ctypes.data_as(vp), newP.ctypes.data_as(vp), rows.ctypes.data_as(vp),
                         np.ascontiguousarray(syms[2 * warm:]).ctypes.data_as(vp), sub, stats.ctypes.data_as(vp))
        for t in range(1, 9):
            o.update_blk(syms[2 * (warm + t - 1):], 1)
            mt = o.get_metrics().astype(np.int64)
            assert int(stats[t]) + sub - 32768 == int(mt[0]), f"state-0 metric after stage {t}"
            assert int(stats[9 + t]) + sub - 32768 == int(mt.min()), f"global min after stage {t}"
            canon = np.zeros(1 << 18, np.uint32)
            e.emu_canon_row(fmt_base + t, rows[t - 1].ctypes.data_as(vp), canon.ctypes.data_as(vp))
            assert np.array_equal(canon, o.get_row(warm + t - 1)), f"decision row of stage {t}"
        m1 = o.get_metrics()
    assert np.array_equal(newP.astype(np.int64) + sub - 32768, m1.astype(np.int64))
    assert int(stats[18]) == int(newP.max())
