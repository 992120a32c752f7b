"""GPU tier (pytest -m gpu, on the B200 box): the CUDA path, called through the C ABI, against
  * the golden fixtures recorded from the unmodified reference,
  * the CPU checker (the reference itself if oracle/_ref travelled, else the oracle port) on fresh seeded inputs,
  * size-independent properties at BASELINE.json's full sizes (known-answer decode, idempotence,
    block == per-bit streaming, checkpoint-window continuation).
Bar: bit-exact (integer arithmetic): decoded bytes, every decision row, every path metric,
renormalisation counts and min/max metrics."""
import os
import subprocess
import sys

import numpy as np
import pytest

import isee3_decoder_b200 as v224
import pyoracle
from scripts import run_script, compare_outcomes, crc

pytestmark = pytest.mark.gpu
S = v224.streams
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(f[:-4] for f in os.listdir(os.path.join(os.path.dirname(__file__), "golden")) if f.endswith(".npz"))


def gpu_factory(**opts):
    def make(n):
        d = v224.Viterbi224(n)
        for k, val in opts.items():
            d.set_option(k, val)
        return d
    return make


@pytest.fixture(scope="module", autouse=True)
def need_gpu(built):
    # fail loudly rather than skip: -m gpu on a box without the CUDA path is a broken run
    assert v224.device_count() > 0, "no CUDA device visible: the CUDA path cannot be exercised"


# ---------------------------------------------------------------------------------------------
# golden fixtures (unmodified reference), every kernel variant
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN)
def test_golden_default_path(golden_cases, name):
    case = golden_cases[name]
    got = run_script(gpu_factory(), case["script"], case["syms"])
    compare_outcomes(got, case["outcome"], f"gpu vs golden {name}")


@pytest.mark.parametrize("opts", [{"force_single": 1}, {"force_sat": 1}, {"force_careful": 1}, {"per_pass_launch": 1},
                                  {"per_pass_launch": 1, "force_careful": 1},
                                  {"chain_seg": 8, "chain_warm": 0},
                                  {"chain_seg": 64, "chain_warm": 16},
                                  {"tile32": 1}, {"tile32": 0}, {"tile32": 1, "force_careful": 1}, {"tile32": 1, "per_pass_launch": 1},
                                  {"tile32": 1, "grid_limit": 97}, {"no_mailbox": 1}, {"slow_single": 1},
                                  {"tile32": 2}, {"no_discard": 1}, {"tile32": 0, "grid_limit": 444}, {"measure_all": 1}])
@pytest.mark.parametrize("name", ["awgn1db_648_chunks", "erasure_304", "ringwrap_len40_200", "saturation_forced_72", "stream_d64_320"])
def test_golden_kernel_variants(golden_cases, name, opts):
    case = golden_cases[name]
    got = run_script(gpu_factory(**opts), case["script"], case["syms"])
    compare_outcomes(got, case["outcome"], f"gpu{opts} vs golden {name}")


def test_golden_uses_the_fused_kernel(golden_cases):
    case = golden_cases["awgn1db_648_chunks"]
    with v224.Viterbi224(648) as d:
        d.init(0)
        d.update_blk(case["syms"], 648)
        st = d.stats()
    assert st["fused_passes"] == 81 and st["single_stages"] == 0 and st["sat_stages"] == 0, st
    assert st["careful_passes"] >= 2, st           # two renormalisations happen in this stream


def test_saturating_fallback_engages(golden_cases):
    case = golden_cases["saturation_forced_72"]
    got = run_script(gpu_factory(), case["script"], case["syms"])
    compare_outcomes(got, case["outcome"], "saturation")
    with v224.Viterbi224(80) as d:
        m = np.random.default_rng(10).integers(27000, 32701, 1 << 23).astype(np.int16)
        m[0] = 15000
        m[1 << 22] = 15000
        d.set_state(m, 0, 0)
        d.update_blk(case["syms"], 72)
        st = d.stats()
    assert st["sat_stages"] >= 2, st               # stages where the reference's adds clip ran in the exact kernel
    assert st["fused_passes"] >= 1, st             # and the decoder returned to the fused kernel afterwards


# ---------------------------------------------------------------------------------------------
# fresh inputs against the CPU checker
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ebn0,style,seed", [(1.0, "vtest", 31), (3.0, "symdemod", 32), (-1.0, "vtest", 33)])
def test_frame_decode_equals_cpu_checker(ebn0, style, seed):
    n = 328                                       # > one renormalisation period at 1 dB, not a multiple of 8*k
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, n, dtype=np.uint8)
    bits[-24:] = 0
    sym01, _ = S.encode_bits(bits, 0)
    syms = S.awgn_vtest(sym01, ebn0, rng) if style == "vtest" else S.awgn_symdemod(sym01, ebn0, rng)
    script = [["create", n], ["init", 0], ["update", 0, 200], ["minmax"], ["update", 200, 128], ["minmax"],
              ["chainback", n, 0], ["decodebit", 100, -1], ["decodeword", 64, 0]]
    Checker = pyoracle.best_cpu_decoder()
    want = run_script(lambda k: Checker(k), script, syms)
    got = run_script(gpu_factory(), script, syms)
    compare_outcomes(got, want, f"gpu vs {Checker.kind}")


def test_empty_and_degenerate_calls():
    with v224.Viterbi224(16) as d:
        assert d.update_blk(np.zeros(0, np.uint8), 0) == 0
        assert d.chainback(0, 0).size == 0
        assert d.decodebit(0, 0) == -1 and d.decodebit(-3, 0) == -1
        assert d.init(0x7FFFFFFF) == 0                      # start state is masked to 23 bits (viterbi224_sse2.c:50)
        m = d.get_metrics()
        assert m[0x7FFFFF] == -32768 and m[0] == -32768 + 5000
        assert d.min_metric() == -32768 and d.max_metric() == -32768 + 5000
        # one single stage, one ragged block, ring wrap at len 16
        syms = np.random.default_rng(1).integers(0, 256, 2 * 45, dtype=np.uint8)
        with pyoracle.best_cpu_decoder()(16) as o:
            o.init(0x7FFFFFFF)
            for a, n in [(0, 1), (1, 17), (18, 27)]:
                assert d.update_blk(syms[2 * a:], n) == o.update_blk(syms[2 * a:], n)
            assert np.array_equal(d.get_metrics(), o.get_metrics())
            for r in range(16):
                assert crc(d.get_row(r)) == crc(o.get_row(r)), r
            assert d.decodebit(15, 0) == o.decodebit(15, 0)
            assert np.array_equal(d.chainback(45, 9), o.chainback(45, 9))


def test_recycled_decoder_is_indistinguishable_from_a_fresh_one(golden_cases):
    """delete_viterbi224 parks the decoder, the next create_viterbi224 of the same size gets it back (decode.c:216-229
    creates and deletes per frame): default options, zero counters, the rows it wrote read as zero, init(0) state --
    and a golden script runs on it exactly as on a new one."""
    syms = np.random.default_rng(3).integers(0, 256, 2 * 100, dtype=np.uint8)
    d = v224.Viterbi224(72)
    d.set_option("force_careful", 1)
    d.init(0x1234)
    d.update_blk(syms, 100)                              # wraps the 72-row ring
    assert any(crc(d.get_row(k)) != crc(np.zeros(1 << 18, np.uint32)) for k in range(72))
    d.delete()
    with v224.Viterbi224(72) as r:
        st = r.stats()
        assert st["stages"] == 0 and st["fused_passes"] == 0 and st["launches"] <= 2 and st["renormals"] == 0, st
        zero = crc(np.zeros(1 << 18, np.uint32))
        assert all(crc(r.get_row(k)) == zero for k in range(72))
        m = r.get_metrics()
        assert m[0] == -32768 and m[1] == -32768 + 5000 and m[0x1234] == -32768 + 5000
    case = golden_cases["saturation_forced_72"]
    got = run_script(gpu_factory(), case["script"], case["syms"])      # the suite as a whole creates and deletes same-sized decoders all the time
    compare_outcomes(got, case["outcome"], "recycled decoder vs golden")


def test_block_stream_equals_per_bit_abi_loop():
    """v224x_stream_decode == the vdecode.c:145-152 loop through the nine-entry ABI (both on the GPU)."""
    bits, syms = S.telemetry_stream(700, 3.0, seed=41)
    delay = 96
    with v224.Viterbi224(delay + 1) as d:                   # vdecode.c:94
        d.init(0)
        per_bit = np.empty(700, np.uint8)
        ren = 0
        for i in range(700):
            ren += d.update_blk(syms[2 * i: 2 * i + 2], 1)
            per_bit[i] = d.decodebit(delay, 0)
        m1 = d.get_metrics()
    with v224.Viterbi224(delay + 256) as d:                 # several chunks of 256
        d.init(0)
        blk, ren2 = d.stream_decode(syms, delay)
        m2 = d.get_metrics()
    assert ren == ren2
    assert np.array_equal(m1, m2)
    # the first `delay` outputs walk past the start of the stream; vdecode suppresses them (vdecode.c:151-158)
    assert np.array_equal(per_bit[delay:], blk[delay:])
    lag = delay + 22
    assert np.array_equal(blk[lag:], bits[:700 - lag])       # and they are the transmitted data


def test_lockstep_multi_decoder_update_equals_separate_updates():
    """v224x_update_multi_dev: three decoders with different streams (one of them mid-frame, one with a ragged
    length) advanced by one persistent launch per batch == each advanced alone == the CPU checker."""
    n = 203                                      # 25 fused passes + 3 single stages
    streams = [S.vtest_frame(208, 1.0, seed=61)[1], S.telemetry_stream(208, 2.0, seed=62)[1], S.vtest_frame(208, 4.0, seed=63)[1]]
    decs = [v224.Viterbi224(64 + 16 * i) for i in range(3)]
    try:
        dptr = []
        for d, sy in zip(decs, streams):
            p = d.dev_alloc(sy.size)
            d.h2d(p, sy)
            dptr.append(p)
        decs[1].init(0x12345)
        decs[1].update_blk(streams[1][:20], 10)          # decoder 1 starts the lockstep call mid-stream
        want = []
        for i, (d, sy) in enumerate(zip(decs, streams)):
            with pyoracle.best_cpu_decoder()(d.len) as o:
                if i == 1:
                    o.init(0x12345)
                    o.update_blk(sy[:20], 10)
                r = o.update_blk(sy, n)
                want.append((r, o.get_metrics(), [crc(o.get_row(k)) for k in range(d.len)], o.min_metric(), o.max_metric()))
        ren = v224.Viterbi224.update_multi_dev(decs, dptr, n)
        for i, d in enumerate(decs):
            assert ren[i] == want[i][0]
            assert np.array_equal(d.get_metrics(), want[i][1])
            assert [crc(d.get_row(k)) for k in range(d.len)] == want[i][2]
            assert (d.min_metric(), d.max_metric()) == (want[i][3], want[i][4])
        assert decs[0].stats()["fused_passes"] == 25
    finally:
        for d in decs:
            d.delete()


def test_per_bit_decodebit_stops_where_it_rejoins_the_previous_walk():
    """The vdecode.c:145-152 pattern (update(1) + decodebit(delay, 0) per bit): the incremental walk returns what the
    full walk returns, with a small fraction of the dependent ring loads; chunked updates and a changed delay fall back."""
    n, delay = 600, 200
    bits, syms = S.telemetry_stream(n, 3.0, seed=43)
    outs = []
    for opt in (1, 0):
        with v224.Viterbi224(delay + 1) as d:                      # vdecode.c:94
            d.set_option("no_walk_cache", opt)
            d.init(0)
            o = []
            for i in range(n):
                d.update_blk(syms[2 * i: 2 * i + 2], 1)
                o.append(d.decodebit(delay, 0))
                if i == 300:
                    o.append(d.decodebit(delay // 2, 0))           # another delay: full walk, then back
                    o.append(d.decodebit(delay, 5))                # another end state
            d.update_blk(syms[:14], 7)                             # several rows at once, then one walk
            o.append(d.decodebit(delay, 0))
            steps = d.stats()["walk_steps"]
        outs.append((o, steps))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == 0                                         # the plain walk does not count
    assert 0 < outs[1][1] < (n + 3) * delay // 3, outs[1][1]       # far fewer than delay loads per call


# ---------------------------------------------------------------------------------------------
# the reference's own programs on our library (drop-in), and the vdecode mirror
# ---------------------------------------------------------------------------------------------
def _bin(name):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not built (reference checkout absent at build time)")
    return p


def test_stock_vtest224_runs_on_the_gpu_library():
    out = subprocess.run([_bin("vtest224_b200"), "-e", "3", "-l", "4096", "-n", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "BER 0/8192" in out.stdout and "FER 0/2" in out.stdout, out.stdout


def test_stock_vdecode_on_gpu_library_and_block_mirror_agree_with_reference_output():
    """One stream with an odd junk prefix (forces vdecode's phase flip, vdecode.c:126-133):
    stock vdecode.c + reference decoder (fixture)  ==  stock vdecode.c + our library  ==  block-mode mirror."""
    bits, soft = S.telemetry_stream(3 * 1024, 6.0, seed=11, junk_symbols=101)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "vdecode_flip_seed11.npy"))
    mirror = v224.vdecode.vdecode(lambda n: v224.Viterbi224(n), soft, delay=64)
    assert np.array_equal(mirror, fx)
    out = subprocess.run([_bin("vdecode_b200"), "-d", "64", "-q"], input=soft.tobytes(), capture_output=True, timeout=900)
    assert out.returncode == 0, out.stderr
    assert np.array_equal(np.frombuffer(out.stdout, dtype=np.uint8), fx)


def _status_lines(stderr_bytes):
    """stderr without the program name (argv[0] differs between the two programs)."""
    return [ln.split(b": ", 1)[1] for ln in stderr_bytes.splitlines() if b": " in ln]


def test_native_block_driver_prints_what_stock_vdecode_prints():
    """isee3-decoder_b200/bin/vdecode_block (C++ host program, one block call per 20,000 pairs, 3 decoders in lockstep)
    against stock vdecode.c on the same library and against the fixture recorded from the unmodified reference:
    stdout byte for byte, stderr (flip notices, re-encode symbol-error tally per status interval) line for line."""
    blk = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    assert os.path.exists(blk), "vdecode_block not built (python __graft_entry__.py build)"
    # (1) the reference's own output (fixture), small stream with a phase flip
    bits, soft = S.telemetry_stream(3 * 1024, 6.0, seed=11, junk_symbols=101)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "vdecode_flip_seed11.npy"))
    out = subprocess.run([blk, "-d", "64", "-q"], input=soft.tobytes(), capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert np.array_equal(np.frombuffer(out.stdout, dtype=np.uint8), fx)
    # (2) a longer stream (several blocks, segmented decode inside, two flips: junk prefix + one dropped symbol mid-stream)
    bits, soft = S.telemetry_stream(46 * 1024, 3.0, seed=12, junk_symbols=33)
    soft = np.delete(soft, 61_007)
    env = dict(os.environ, LANG="C")
    args = ["-d", "200", "-i", "5000"]
    a = subprocess.run([_bin("vdecode_b200")] + args, input=soft.tobytes(), capture_output=True, timeout=900, env=env)
    b = subprocess.run([blk] + args + ["-B", "20000"], input=soft.tobytes(), capture_output=True, timeout=300, env=env)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-300:], b.stderr[-300:])
    assert a.stdout == b.stdout and len(a.stdout) > 46_000
    la, lb = _status_lines(a.stderr), _status_lines(b.stderr)
    assert la == lb and sum(b"flipping phase" in x for x in la) >= 2 and sum(b"symerrs" in x for x in la) >= 9, la


def _strip_argv0(stdout_bytes):
    """decode prints argv[0] in front of its banner lines; everything else is compared byte for byte."""
    return [ln.split(b": ", 1)[1] if (b": Fano" in ln or b": Not displaying" in ln) else ln for ln in stdout_bytes.split(b"\n")]


def test_frame_decoder_prints_what_stock_decode_prints():
    """isee3-decoder_b200/bin/decode_block (frames of a locked run decoded side by side, speculatively, in one
    v224x_decode_frames call) against `decode -V`:
    (1) the output recorded from the UNMODIFIED reference (decode.c + viterbi224_sse2.c, tools/make_golden_host.py):
        junk prefix, false first sync (bad frame), lock, 100 symbols lost (bad frame, re-acquisition), 5 symbols inserted;
    (2) the same for 60 frames at 3 dB with four disturbances (56 frames, 5 bad);
    (3) stock decode.c linked against our library on that stream, with and without -n, at several batch sizes."""
    blk = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
    assert os.path.exists(blk), "decode_block not built (python __graft_entry__.py build)"
    env = dict(os.environ, LANG="C", V224_HOST_STATS="1")
    fx = np.load(os.path.join(ROOT, "tests", "golden", "host", "decode_V_seed21.npz"))
    want = _strip_argv0(bytes(fx["stdout"]))
    for extra in ([], ["-B", "1"], ["-B", "3", "-L", "2"]):
        out = subprocess.run([blk, "-V"] + extra, input=fx["symbols"].tobytes(), capture_output=True, timeout=300, env=env)
        assert out.returncode == 0, out.stderr
        assert _strip_argv0(out.stdout) == want, extra
    assert sum(ln.startswith(b"Frame ") for ln in want) == 7 and sum(ln.endswith(b"(bad)") for ln in want) == 3
    # (2) 60 frames at 3 dB, symbols cut (40, 1000) and inserted (7, 300) inside four different frames, truncated end:
    #     the UNMODIFIED reference printed 56 frames, 5 of them bad (fixture; ~2.5 min of CPU when it was recorded)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import zlib
    import make_golden_host
    soft = make_golden_host.decode_stream_long()
    fxl = np.load(os.path.join(ROOT, "tests", "golden", "host", "decode_V_long_seed77.npz"))
    assert zlib.crc32(soft.tobytes()) == int(fxl["symbols_crc"]), "stream generator drifted from the fixture"
    out = subprocess.run([blk, "-V"], input=soft.tobytes(), capture_output=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    assert _strip_argv0(out.stdout) == _strip_argv0(bytes(fxl["stdout"]))
    assert out.stdout.count(b"Frame ") == 56 and out.stdout.count(b"(bad)") == 5
    # (3) the same stream through stock decode.c linked against our library, with and without -n, at another batch size
    for flags in (["-V"], ["-V", "-n", "-r", "2048"]):
        a = subprocess.run([_bin("decode_b200")] + flags, input=soft.tobytes(), capture_output=True, timeout=900, env=env)
        assert a.returncode == 0, a.stderr[-300:]
        for extra in ([], ["-B", "5"]):
            b = subprocess.run([blk] + flags + extra, input=soft.tobytes(), capture_output=True, timeout=300, env=env)
            assert b.returncode == 0, b.stderr[-300:]
            assert _strip_argv0(a.stdout) == _strip_argv0(b.stdout), (flags, extra)
        assert a.stdout.count(b"Frame ") == (51 if "-n" in flags else 56)
    assert b"speculative frames discarded" in b.stderr


def test_frame_decoder_fano_first_mode_prints_what_stock_decode_prints():
    """decode_block in the reference's default mode -- Fano on the host first, the frames it gives up on decoded by the GPU
    in one batch per run -- against the printouts recorded from the UNMODIFIED reference (`decode`, `decode -p -n`;
    tools/make_golden_host.py): which decoder each frame is credited to, partial Fano output of frames that were not
    retried, lock losses; and against stock decode.c on our library for a longer stream."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import zlib
    import make_golden_host
    blk = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
    env = dict(os.environ, LANG="C", V224_HOST_STATS="1")
    n_viterbi = 0
    for name, flags, soft in make_golden_host.hybrid_cases():
        fx = np.load(os.path.join(ROOT, "tests", "golden", "host", name + ".npz"))
        assert zlib.crc32(soft.tobytes()) == int(fx["symbols_crc"]), "stream generator drifted from the fixture"
        for extra in ([], ["-B", "3"]):
            out = subprocess.run([blk] + flags + extra, input=soft.tobytes(), capture_output=True, timeout=300, env=env)
            assert out.returncode == 0, out.stderr
            assert _strip_argv0(out.stdout) == _strip_argv0(bytes(fx["stdout"])), (name, extra)
        n_viterbi += out.stdout.count(b"with Viterbi")
    assert n_viterbi >= 20
    _, soft = S.telemetry_stream(120 * 1024, 1.9, seed=31, junk_symbols=700)
    for flags in ([], ["-p"]):
        a = subprocess.run([_bin("decode_b200")] + flags, input=soft.tobytes(), capture_output=True, timeout=900, env=env)
        b = subprocess.run([blk] + flags, input=soft.tobytes(), capture_output=True, timeout=300, env=env)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr[-300:], b.stderr[-300:])
        assert _strip_argv0(a.stdout) == _strip_argv0(b.stdout), flags
        assert a.stdout.count(b"with Viterbi") >= 3 and a.stdout.count(b"with Fano") >= 80


def test_framing_mode_equals_the_vdecode_framer_pipeline():
    """vdecode_block -f prints what `vdecode | framer` prints (framer.c:61-95): here vdecode_block's own bit stream piped
    through the unmodified reference framer (oracle/_ref/framer_ref), on a stream with a phase flip."""
    blk = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    env = dict(os.environ, LANG="C")
    _, soft = S.telemetry_stream(24 * 1024, 3.0, seed=14, junk_symbols=55)
    bits = subprocess.run([blk, "-d", "200", "-q"], input=soft.tobytes(), capture_output=True, timeout=300, env=env)
    assert bits.returncode == 0, bits.stderr
    want = subprocess.run([_bin("framer_ref"), "-r", "512"], input=bits.stdout, capture_output=True, timeout=60, env=env)
    got = subprocess.run([blk, "-d", "200", "-q", "-f", "-r", "512", "-B", "7000"], input=soft.tobytes(), capture_output=True, timeout=300, env=env)
    assert got.returncode == 0, got.stderr
    assert got.stdout == want.stdout and got.stdout.count(b"Frame ") >= 20


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------
def test_config1_full_size_known_answer():
    """vtest224 -l 10000 -e 3: create(10000) / init / update / chainback, BER 0 (vtest224.c:100-130)."""
    data, syms = S.vtest_frame(10000, 3.0, seed=1001)
    with v224.Viterbi224(10000) as d:
        d.init(0)
        d.update_blk(syms, 10000)
        out = d.chainback(10000, 0)
        assert np.array_equal(out, data)
        out2 = d.chainback(10000, 0)                 # traceback is idempotent
        assert np.array_equal(out, out2)
        d.set_option("chain_seg", 10000)             # one serial segment == speculative segments
        assert np.array_equal(d.chainback(10000, 0), out)
        assert d.stats()["fused_passes"] == 1250


def test_long_delay_stream_and_checkpoint_window():
    """Config 3/4 in miniature: long stream at 2 dB, delay 2048, then the SURVEY 8c window check --
    dump the GPU state mid-stream, continue 40 stages on the CPU checker and on the GPU, compare everything."""
    n = 40000
    rng = np.random.default_rng(77)
    bits = rng.integers(0, 2, n, dtype=np.uint8)
    sym01, _ = S.encode_bits(bits, 0)
    syms = S.awgn_vtest(sym01, 2.0, rng)
    delay = 2048
    with v224.Viterbi224(delay + 4096) as d:
        d.init(0)
        out, ren = d.stream_decode(syms, delay)
        lag = delay + 22
        errs = int((out[lag:] != bits[:n - lag]).sum())
        assert errs == 0, errs                       # 2 dB is far above this code's threshold
        assert ren >= n // 400                       # renormalisations happened (about one per 280 bits at these levels)
        st = d.stats()
        assert st["sat_stages"] == 0
        m = d.get_metrics()
        renormals = st["renormals"]
        # continue on both sides from the dumped state
        tail_bits = rng.integers(0, 2, 40, dtype=np.uint8)
        tsym01, _ = S.encode_bits(tail_bits, int("".join(map(str, bits[-24:])), 2))
        tsyms = S.awgn_vtest(tsym01, 2.0, rng)
        with pyoracle.best_cpu_decoder()(delay + 4096) as o:
            o.set_state(m, renormals, n)
            r_cpu = o.update_blk(tsyms, 40)
            r_gpu = d.update_blk(tsyms, 40)
            assert r_cpu == r_gpu
            assert np.array_equal(o.get_metrics(), d.get_metrics())
            assert o.min_metric() == d.min_metric() and o.max_metric() == d.max_metric()
            for k in range(40):
                row = (n + k) % (delay + 4096)
                assert crc(o.get_row(row)) == crc(d.get_row(row)), k


@pytest.mark.parametrize("nseg,style,ebn0", [(2, "telemetry", 3.0), (3, "vtest", 1.0), (4, "telemetry", 2.0)])
def test_lockstep_segmented_stream_equals_sequential(nseg, style, ebn0):
    """v224x_stream_decode_seg: nseg decoders in lockstep over contiguous segments, every hand-over verified on the
    device (metric vectors equal up to a constant) -> output identical to the sequential block decode."""
    n, delay, conv = 70001, 200, 1024
    if style == "telemetry":
        bits, syms = S.telemetry_stream(n, ebn0, seed=90 + nseg)
    else:
        rng = np.random.default_rng(90 + nseg)
        bits = rng.integers(0, 2, n, dtype=np.uint8)
        sym01, _ = S.encode_bits(bits, 0)
        syms = S.awgn_vtest(sym01, ebn0, rng)
    with v224.Viterbi224(delay + 4096) as d:
        d.init(0)
        want, _ = d.stream_decode(syms, delay)
    with v224.Viterbi224(delay + 4096) as d:
        d.init(0)
        got, rep = d.stream_decode_seg(syms, delay, nseg, conv)
        assert rep["segments"] == nseg and rep["verified"] == nseg - 1 and rep["redone"] == 0 and rep["worst_spread"] == 0, rep
        assert rep["warm"] == delay + conv and rep["extra_stages"] == (nseg - 1) * (delay + conv)
        assert np.array_equal(got, want)
        # the handle continues the stream exactly: 40 more stages, per-bit ABI, against a sequential decoder
        tail = syms[:80]
        a = [(d.update_blk(tail[2 * i: 2 * i + 2], 1), d.decodebit(delay, 0))[1] for i in range(40)]
    with v224.Viterbi224(delay + 4096) as d:
        d.init(0)
        d.stream_decode(syms, delay)
        b = [(d.update_blk(tail[2 * i: 2 * i + 2], 1), d.decodebit(delay, 0))[1] for i in range(40)]
    assert a == b
    lag = delay + 22
    if ebn0 >= 3.0:
        assert np.array_equal(got[lag:], bits[: n - lag])


def test_segmented_stream_keeps_handle_identity():
    """The handle's body is exchanged with the last segment's decoder; its stream, timer events and options stay:
    a timer started before an odd number of segmented calls stops cleanly, and the next ABI call works."""
    n, delay = 40000, 64
    bits, syms = S.telemetry_stream(n, 4.0, seed=98)
    with v224.Viterbi224(delay + 2048) as d:
        d.set_option("chain_seg", 64)
        d.init(0)
        d.timer_start()
        got, rep = d.stream_decode_seg(syms, delay, 3, 512)
        ms = d.timer_stop_ms()
        assert rep["segments"] == 3 and rep["redone"] == 0
        assert ms > 0
        assert d.init(0) == 0
        again, _ = d.stream_decode(syms, delay)
        assert np.array_equal(got, again)


def test_segmented_stream_failed_handover_is_redone_exactly():
    """conv = 0 puts the hand-over check at the very start of the later decoders (uniform metrics against the
    true ones): it must fail, and the call must fall back to the sequential decode -- still bit-exact."""
    n, delay = 30000, 96
    bits, syms = S.telemetry_stream(n, 3.0, seed=97)
    with v224.Viterbi224(delay + 2048) as d:
        d.init(0)
        want, _ = d.stream_decode(syms, delay)
    with v224.Viterbi224(delay + 2048) as d:
        d.init(0)
        got, rep = d.stream_decode_seg(syms, delay, 3, 0)
        assert rep["segments"] == 3 and rep["verified"] == 0 and rep["redone"] == 2 and rep["worst_spread"] > 0, rep
        assert np.array_equal(got, want)
        short, rep2 = d.stream_decode_seg(syms[:2 * 3000], delay, 4, 1024)       # too short for segments
        assert rep2["segments"] == 1
    with v224.Viterbi224(delay + 2048) as d:
        d.init(0)
        d.stream_decode(syms, delay)
        want2, _ = d.stream_decode(syms[:2 * 3000], delay)
    assert np.array_equal(short, want2)


def test_batched_frames_equal_the_three_call_sequence():
    """v224x_decode_frames: 7 independent 1024-bit frames (decode.c:220-222 pattern: known start and end state = the
    sync word's low 24 bits; two frames at 1 dB, one with a wrong end state) decoded 3 side by side == each frame through
    init / update / chainback on one decoder == the transmitted data where the channel allows."""
    nframes, fb = 7, 1024
    sync = v224.streams.SYNCWORD & 0xFFFFFF
    rng = np.random.default_rng(606)
    frames = v224.streams.telemetry_bits(nframes, rng).reshape(nframes, fb)
    syms, starts, ends = [], [], []
    state = sync                                         # the previous frame ended with the sync word
    for f in range(nframes):
        sym01, state_out = S.encode_bits(frames[f], state)
        ebn0 = 1.0 if f in (2, 5) else 3.0
        syms.append(S.awgn_symdemod(sym01, ebn0, rng))
        starts.append(state & 0x7FFFFF)
        ends.append(sync if f != 4 else 12345)
        state = state_out
    syms = np.concatenate(syms)
    want = []
    with v224.Viterbi224(fb) as d:
        for f in range(nframes):
            d.init(starts[f])
            d.update_blk(syms[2 * fb * f:], fb)
            want.append(d.chainback(fb, ends[f]))
    with v224.Viterbi224(fb) as d:
        got = d.decode_frames(syms, nframes, fb, starts, ends, nlock=3)
        got1 = d.decode_frames(syms, nframes, fb, starts, ends, nlock=1)
    for f in range(nframes):
        assert np.array_equal(got[f], want[f]), f
        assert np.array_equal(got1[f], want[f]), f
        if f not in (2, 4, 5):
            assert np.array_equal(got[f], np.packbits(frames[f])), f


def _window_check(d, syms_tail, nstages, ring_rows, label):
    """SURVEY 8c checkpoint window: dump the GPU state, continue `nstages` on the CPU checker and on the GPU,
    compare renormalisation counts, every metric and every decision row."""
    m = d.get_metrics()
    st = d.stats()
    T = st["stages"]
    with pyoracle.best_cpu_decoder()(ring_rows) as o:
        o.set_state(m, st["renormals"], T)
        assert o.update_blk(syms_tail, nstages) == d.update_blk(syms_tail, nstages), label
        assert np.array_equal(o.get_metrics(), d.get_metrics()), label
        for k in range(nstages):
            row = (T + k) % ring_rows
            assert crc(o.get_row(row)) == crc(d.get_row(row)), (label, k)


def test_config3_full_size_long_delay_traceback():
    """BASELINE config 3 at full size: 4,194,304 bits at Eb/N0 = 2 dB, decode delay 2048.  Known-answer (decoded ==
    transmitted), segmented == sequential, and a checkpoint window against the CPU checker at the end of the stream."""
    n, delay = 1 << 22, 2048
    rng = np.random.default_rng(303)
    bits = rng.integers(0, 2, n + 48, dtype=np.uint8)
    sym01, _ = S.encode_bits(bits, 0)
    syms = S.awgn_vtest(sym01, 2.0, rng)
    ring = delay + 8192
    with v224.Viterbi224(ring) as d:
        d.init(0)
        seq, _ = d.stream_decode(syms[:2 * n], delay)
    with v224.Viterbi224(ring) as d:
        d.init(0)
        out, rep = d.stream_decode_seg(syms[:2 * n], delay, 3)
        assert rep["segments"] == 3 and rep["verified"] == 2 and rep["redone"] == 0, rep
        assert np.array_equal(out, seq)
        lag = delay + 22
        errs = int((out[lag:] != bits[:n - lag]).sum())
        print(f"config 3: {errs} bit errors in {n - lag} bits at 2 dB (BER {errs / (n - lag):.2e})")
        assert errs / (n - lag) < 1e-4                   # the code's own error rate at 2 dB (8-bit quantised), not a decoder defect
        _window_check(d, syms[2 * n:], 48, ring, "config 3 end of stream")


def test_config4_full_size_low_snr_stress():
    """BASELINE config 4 at full size: 16,777,216 bits at Eb/N0 = 1 dB.  BER against the transmitted data, and
    bit-exactness against the CPU checker through checkpoint windows at three stream offsets (a CPU decode of the whole
    stream would take half a day)."""
    n, delay, parts = 1 << 24, 200, 3
    rng = np.random.default_rng(404)
    bits = rng.integers(0, 2, n, dtype=np.uint8)
    sym01, _ = S.encode_bits(bits, 0)
    syms = S.awgn_vtest(sym01, 1.0, rng)
    ring = delay + 8192
    out = np.full(n, 255, np.uint8)                  # 255 = not produced by the block calls (the window stages)
    win = 40
    parts = [(0, 5_000_003), (5_000_003 + win, 11_000_017), (11_000_017 + win, n - 64)]
    with v224.Viterbi224(ring) as d:
        d.init(0)
        for a, b in parts:
            o, rep = d.stream_decode_seg(syms[2 * a: 2 * b], delay, 3)
            assert rep["segments"] == 3 and rep["redone"] == 0 and rep["worst_spread"] == 0, rep
            out[a:b] = o
            w = win if b < n - 64 else 64
            _window_check(d, syms[2 * b: 2 * (b + w)], w, ring, f"config 4 offset {b}")
    lag = delay + 22
    bers = []
    for a, b in parts:                                # per part: a misaligned or corrupted part would stand out
        lo = max(a, lag)
        e = int((out[lo:b] != bits[lo - lag:b - lag]).sum())
        bers.append(e / (b - lo))
    print("config 4: BER at 1 dB per part " + ", ".join(f"{x:.3e}" for x in bers))
    # 1 dB is below this code's waterfall (config 3: 2e-5 at 2 dB): the BER is the code's, the decoder is checked bit
    # for bit by the windows above.  Sanity: every part decodes (far from 0.5) and the parts agree with each other.
    assert max(bers) < 0.05 and max(bers) < 1.5 * min(bers) + 1e-4


def test_time_segmented_decode_matches_single_pass():
    """Multi-GPU partitioning run on one GPU: segments with warm-up reproduce the single-pass output
    wherever survivors merged inside the warm-up; residual differences are counted (north_star)."""
    n = 24000
    bits, syms = S.telemetry_stream(n, 3.0, seed=55)
    delay = 200
    with v224.Viterbi224(delay + 4096) as d:
        d.init(0)
        full, _ = d.stream_decode(syms, delay)
        parts = []
        for seg in v224.segments.plan(n, 4, warmup=1024, delay=delay):
            parts.append(v224.segments.decode_segment(d, syms, seg, delay))
    seg_out = v224.segments.stitch(parts)
    assert seg_out.size == n
    diff = int((seg_out != full).sum())
    assert diff == 0, f"{diff} residual differences vs single pass"


# ---------------------------------------------------------------------------------------------
# round 2: config 2 at scale against the unmodified reference, the multi-GPU range machinery, stock hybridtest
# ---------------------------------------------------------------------------------------------
def _long_vdecode_fixture():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden_vdecode_long as g
    import zlib
    fx = np.load(os.path.join(ROOT, "tests", "golden", "host", "vdecode_long_seed2014.npz"))
    soft = g.stream()
    assert soft.size == int(fx["nsyms"]) and (zlib.crc32(soft.tobytes()) & 0xFFFFFFFF) == int(fx["soft_crc"]), "stream generator drifted"
    want = np.unpackbits(fx["bits_packed"])[: int(fx["nout"])]
    return g, soft, want, str(fx["stderr"])


def test_config2_at_scale_equals_the_unmodified_reference():
    """BASELINE config 2 at scale: 73,577 decoded bits of a symdemod-format stream (junk prefix -> initial phase flip, one
    symbol lost in mid-stream -> second flip) as printed by the UNMODIFIED vdecode.c + viterbi224_sse2.c
    (tools/make_golden_vdecode_long.py, about 4 CPU-minutes of the reference).
      (1) bin/vdecode_block (pairing on the host, one block call per 30,000 pairs, lockstep segments): stdout and stderr
      (2) the library calls bench.py times: v224x_pair_symbols + v224x_stream_decode_seg
      (3) the multi-GPU context over the same pairs (2 ranges; on a one-GPU box both on device 0)"""
    g, soft, want, want_err = _long_vdecode_fixture()
    blk = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    env = dict(os.environ, LANG="C")
    out = subprocess.run([blk, "-d", str(g.DELAY), "-i", str(g.INTERVAL), "-B", "30000", "-S", "3"], input=soft.tobytes(), capture_output=True,
                         timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-400:]
    got = np.frombuffer(out.stdout, dtype=np.uint8) - ord("0")
    assert got.size == want.size and np.array_equal(got, want), f"{int((got[:want.size] != want[:got.size]).sum())} bits differ from the reference's output"
    assert _status_lines(out.stderr) == [ln.split(": ", 1)[1].encode() for ln in want_err.splitlines() if ": " in ln]
    # (1b) the same program over two GPUs (-G 2; on a one-GPU box both ranges on device 0), blocks of 40,000 pairs: same stdout, and
    #      the hand-over summary says every GPU-to-GPU hand-over was verified
    ndev = v224.device_count()
    out2 = subprocess.run([blk, "-d", str(g.DELAY), "-q", "-B", "40000", "-S", "2", "-G", "2", "-D", f"0,{1 % ndev}", "-v"], input=soft.tobytes(),
                          capture_output=True, timeout=600, env=env)
    assert out2.returncode == 0, out2.stderr[-400:]
    assert out2.stdout == out.stdout
    assert b"GPU-to-GPU hand-overs verified 2, ranges redone 0" in out2.stderr and b"worst snapshot spread 0" in out2.stderr, out2.stderr[-300:]
    # (2)
    pairs, flips = v224.pair_symbols(soft)
    assert len(flips) == 2
    n = pairs.shape[0]
    with v224.Viterbi224(g.DELAY + 8192) as d:
        d.init(0)
        bits, rep = d.stream_decode_seg(pairs.reshape(-1), g.DELAY, 3)
        assert rep["segments"] == 3 and rep["redone"] == 0
    assert np.array_equal(bits[g.DELAY:], want)
    # (3)
    with v224.MultiGpu(2, g.DELAY + 4096, devices=[0, 1 % ndev]) as m:
        m.init(0)
        mbits, mrep = m.stream_decode(pairs.reshape(-1), g.DELAY, nseg=2, conv=2048)
    assert mrep["gpus"] == 2 and mrep["handovers_verified"] == 1 and mrep["ranges_redone"] == 0 and mrep["worst_spread"] == 0, mrep
    assert np.array_equal(mbits, bits)


def test_config2_full_size_checkpoint_windows():
    """BASELINE config 2 at full size (1,048,576 bits, symdemod format at 3 dB, odd junk prefix -> phase flip after the first
    sync period, plus whatever false-alarm flips the correlator produces at that SNR -- reference behaviour):
    SURVEY 8c checkpoint windows at three offsets of the pair stream -- GPU state dumped, 48 stages continued on the CPU
    checker and on the GPU, renormalisation counts, every metric and every decision row compared (a CPU decode of the whole
    stream would take 45 minutes) -- the parts equal one uninterrupted decode, and that decode is the transmitted data
    everywhere outside the flip transients."""
    sys.path.insert(0, ROOT)
    import bench
    n, delay, ring, win = 1 << 20, 200, 200 + 8192, 48
    bits, soft = S.telemetry_stream(n, 3.0, seed=20141, junk_symbols=101)
    pairs, flips = v224.pair_symbols(soft)
    assert 1 <= len(flips) <= 9 and flips[0] in (2047, 4095), flips
    syms = pairs.reshape(-1)
    npairs = pairs.shape[0]
    parts = [(0, 300_011), (300_011 + win, 700_001), (700_001 + win, npairs - win)]
    outs = []
    with v224.Viterbi224(ring) as d:
        d.init(0)
        for a, b in parts:
            o, rep = d.stream_decode_seg(syms[2 * a: 2 * b], delay, 3)
            assert rep["segments"] == 3 and rep["redone"] == 0 and rep["worst_spread"] == 0, rep
            outs.append(o)
            _window_check(d, syms[2 * b: 2 * (b + win)], win, ring, f"config 2 offset {b}")
        d.init(0)
        full, _ = d.stream_decode_seg(syms, delay, 3)
    for (a, b), o in zip(parts, outs):
        assert np.array_equal(o, full[a:b]), (a, b)
    errs, nchk, ntrans = bench.ber_check(full, bits, flips, delay, (101 - 1) // 2)
    print(f"config 2: {errs} bit errors in {nchk} bits, {ntrans} bits inside phase-flip transients, flips at pairs {flips}")
    assert errs == 0 and nchk > npairs - 16384 * len(flips)


def test_range_decode_handover_is_verified_and_failure_is_detected():
    """The multi-GPU primitive on one GPU: two handles, range 0 continues from init(0) and leaves a late snapshot, range 1
    starts delay + conv stages early from uniform metrics and leaves an early snapshot; the snapshots differ by a constant
    (spread 0) and the stitched output is the sequential decode.  With conv = 0 the early snapshot is the uniform start
    vector: the spread is large and the caller knows the range must be redone."""
    n, delay, conv = 60000, 200, 1536
    bits, syms = S.telemetry_stream(n, 2.5, seed=401)
    cut = 29_996
    with v224.Viterbi224(delay + 4096) as ref:
        ref.init(0)
        want, _ = ref.stream_decode(syms, delay)
    with v224.Viterbi224(delay + 4096) as a, v224.Viterbi224(delay + 4096) as b:
        snap_a, snap_b = a.dev_alloc(1 << 24), b.dev_alloc(1 << 24)
        a.init(0)
        out_a, rep_a = a.range_decode(syms[: 2 * cut], 0, cut, delay, nseg=2, conv=conv, snap_late=snap_a)
        lead = delay + conv
        out_b, rep_b = b.range_decode(syms[2 * (cut - lead):], lead, n - cut, delay, nseg=2, conv=conv, snap_early=snap_b)
        assert a.metric_spread_dev(snap_a, snap_b) == 0
        assert np.array_equal(np.concatenate([out_a, out_b]), want)
        assert rep_a["segments"] == 2 and rep_b["segments"] == 2
        # the handle of range 1 continues the stream exactly (up to a constant metric offset)
        more, _ = b.stream_decode(syms[:400], delay)
        with v224.Viterbi224(delay + 4096) as c:
            c.init(0)
            c.stream_decode(syms, delay)
            more_want, _ = c.stream_decode(syms[:400], delay)
        assert np.array_equal(more, more_want)
        # conv = 0: checked at the uniform start vector -> not converged
        b.range_decode(syms[2 * (cut - delay):], delay, n - cut, delay, nseg=1, conv=0, snap_early=snap_b)
        assert a.metric_spread_dev(snap_a, snap_b) > 1000
        a.dev_free(snap_a)
        b.dev_free(snap_b)


@pytest.mark.parametrize("ngpu,conv", [(2, 1024), (3, 1024), (2, 0)])
def test_multi_gpu_context_equals_sequential_and_redoes_failed_ranges(ngpu, conv):
    """v224x_multi_*: the stream in ngpu ranges (one host thread per range; the devices are the box's GPUs round-robin, so on a
    one-GPU box the ranges share device 0 and the peer copy degenerates to a pointer), two consecutive block calls on one
    context (the second continues on the GPU that holds the stream's end), output == one sequential decoder.  conv = 0 makes
    every GPU-to-GPU check fail: every later range is decoded again by the previous range's decoder, output still exact."""
    n, delay = 90_000, 128
    bits, syms = S.telemetry_stream(n, 3.0, seed=77 + ngpu)
    with v224.Viterbi224(delay + 4096) as d:
        d.init(0)
        want, _ = d.stream_decode(syms, delay)
    ndev = v224.device_count()
    first = 50_000
    with v224.MultiGpu(ngpu, delay + 4096, devices=[i % ndev for i in range(ngpu)]) as m:
        m.init(0)
        a, rep_a = m.stream_decode(syms[: 2 * first], delay, nseg=2, conv=conv)
        b, rep_b = m.stream_decode(syms[2 * first:], delay, nseg=2, conv=conv)
    got = np.concatenate([a, b])
    assert np.array_equal(got, want), f"{int((got != want).sum())} bits differ from the sequential decode"
    for rep in (rep_a, rep_b):
        assert rep["gpus"] >= 2, rep
        if conv:
            assert rep["handovers_verified"] == rep["gpus"] - 1 and rep["ranges_redone"] == 0 and rep["worst_spread"] == 0, rep
        else:
            assert rep["handovers_verified"] == 0 and rep["ranges_redone"] == rep["gpus"] - 1 and rep["worst_spread"] > 0, rep


def test_short_ring_decodes_exactly():
    """Rings shorter than two fused passes (len < 16): consecutive passes of one persistent launch would share ring rows, so
    such handles run one launch per pass.  Frame decode and per-bit streaming on len = 9 and 12 against the CPU checker."""
    Checker = pyoracle.best_cpu_decoder()
    for length, seed in [(9, 1), (12, 2), (15, 3)]:
        bits, syms = S.telemetry_stream(96, 4.0, seed=500 + seed)
        script = [["create", length], ["init", 0], ["update", 0, 40], ["decodebit", length - 1, 0], ["update", 40, 17], ["minmax"],
                  ["decodebit", 5, -1], ["update", 57, 39], ["decodeword", length - 1, 0]]
        got = run_script(gpu_factory(), script, syms)
        ref = run_script(lambda n: Checker(n), script, syms)
        compare_outcomes(got, ref, f"short ring len {length}")


def test_stock_hybridtest_runs_on_the_gpu_library():
    """hybridtest.c compiled unchanged and linked against libviterbi224_b200.so (oracle/_ref/hybridtest_b200), seed pinned by
    the preloaded time() shim: 30 frames at 1 dB, 15 of them fall through to the Viterbi decoder (create / init / update /
    chainback / delete per frame, hybridtest.c:186-193).  Every line it prints -- Fano cycle counts, per-frame Viterbi verdicts
    with their bit-error counts, the two summary lines -- equals the printout of the same program on the reference's SSE2
    decoder (tests/golden/host/hybridtest_seed20141_1dB.txt, tools/make_golden_hybridtest.py)."""
    shim = os.path.join(ROOT, "oracle", "_ref", "libfixed_time.so")
    if not os.path.exists(shim):
        pytest.skip("oracle/_ref/libfixed_time.so not built (reference checkout absent at build time)")
    env = dict(os.environ, LD_PRELOAD=shim, V224_FIXED_TIME="20141")
    out = subprocess.run([_bin("hybridtest_b200"), "-n", "30", "-e", "1.0", "-v"], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-400:]
    want = open(os.path.join(ROOT, "tests", "golden", "host", "hybridtest_seed20141_1dB.txt")).read()

    def lines(t):
        return [ln for ln in t.splitlines() if not ln.startswith("gen_met(")]       # that line prints a stack address
    got_l, want_l = lines(out.stdout), lines(want)
    assert got_l == want_l, [(a, b) for a, b in zip(got_l, want_l) if a != b][:3]
    assert "Viterbi attempts 15 good frames: 11 frame errors 4" in out.stdout


def test_two_host_threads_decode_concurrently_on_one_gpu():
    """The library is used from several host threads at once by the multi-GPU context (one thread per GPU) and may be by callers:
    two threads, each with its own handle on the same GPU, run segmented stream decodes, frame decodes and create / delete cycles
    at the same time; every result equals the single-threaded one (launch geometry cache, tensor-map encoder, decoder and ring
    pools are shared state)."""
    import threading
    n, delay = 40_000, 100
    streams = [S.telemetry_stream(n, 3.0, seed=700 + i)[1] for i in range(2)]
    want = []
    for syms in streams:
        with v224.Viterbi224(delay + 2048) as d:
            d.init(0)
            want.append(d.stream_decode(syms, delay)[0])
    got = [None, None]
    errors = []

    def work(i):
        try:
            for rep in range(3):
                with v224.Viterbi224(delay + 2048) as d:          # create / delete every round: the pools are exercised too
                    d.init(0)
                    out, rep_ = d.stream_decode_seg(streams[i], delay, 3, 1024)
                    assert rep_["redone"] == 0
                    data, fs = S.vtest_frame(512, 3.0, seed=900 + i)
                    d.init(0)
                    d.update_blk(fs, 512)
                    assert np.array_equal(d.chainback(512, 0), data)
                got[i] = out
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for i in range(2):
        assert np.array_equal(got[i], want[i]), i


@pytest.mark.parametrize("name,lo,hi,pins", [
    ("wide spread, fast path but no L2 discard", -32000, -9500, [[0, -20000]]),
    ("spread beyond the fast path: exact integer stages until it narrows", -32768, -5000, [[0, -32768], [77, 2000]]),
    ("close to saturation with a small spread: careful passes and saturating stages", 22500, 24200, [[0, 23000]]),
    ("state 0 just below the renormalisation trigger, minimum far below", -30000, 24000, [[0, 24990], [4194304, 24500]]),
])
def test_synthetic_states_at_the_limits_of_the_fast_path(name, lo, hi, pins):
    """Mid-stream states loaded through v224x_set_state that sit at the edges the fused kernel's bookkeeping watches -- spread near
    MAX_FAST_SPREAD (passes may run but must keep their input: no L2 discard), spread beyond it, metrics near int16 saturation,
    state 0 at the renormalisation trigger -- then 88 stages in chunks that mix fused passes, remainders and the lockstep path.
    Renormalisation counts, min / max metrics, every path metric and every decision row equal the CPU checker."""
    Checker = pyoracle.best_cpu_decoder()
    rng = np.random.default_rng(len(name))
    syms = rng.integers(0, 256, 2 * 96, dtype=np.uint8)
    script = [["create", 96], ["init", 0], ["set_state", 31, lo, hi, pins, 1234, 5], ["update", 0, 16], ["minmax"], ["update", 16, 3],
              ["update", 19, 48], ["minmax"], ["update", 67, 21], ["decodebit", 70, -1], ["chainback", 90, 0]]
    got = run_script(gpu_factory(), script, syms)
    ref = run_script(lambda n: Checker(n), script, syms)
    compare_outcomes(got, ref, name)
    got2 = run_script(gpu_factory(tile32=0, grid_limit=444), script, syms)       # the 64-column build (the one that discards)
    compare_outcomes(got2, ref, name + " / 64-column tiles")


def test_lockstep_decoders_in_limit_states_leave_lockstep_exactly():
    """Four decoders in one lockstep launch (the 64-column build, which drops consumed metric lines from the L2), three of them in
    synthetic states the bookkeeping has to treat specially: near int16 saturation (passes are declined, exact saturating stages
    take over, the decoder leaves lockstep), a spread near the fast path's limit (passes run but keep their input), state 0 at
    the renormalisation trigger.  Each decoder ends exactly where the CPU checker ends on its own stream."""
    Checker = pyoracle.best_cpu_decoder()
    n, ring = 120, 128
    rng = np.random.default_rng(99)
    streams = [rng.integers(0, 256, 2 * n, dtype=np.uint8) for _ in range(4)]
    states = [None, (27000, 32700, [[0, 15000], [4194304, 15000]]), (-32000, -9500, [[0, -20000]]), (-30000, 24000, [[0, 24990]])]
    decs = [v224.Viterbi224(ring) for _ in range(4)]
    try:
        dptr, want = [], []
        for i, (d, sy) in enumerate(zip(decs, streams)):
            p = d.dev_alloc(sy.size)
            d.h2d(p, sy)
            dptr.append(p)
            m = None
            if states[i]:
                lo, hi, pins = states[i]
                m = np.random.default_rng(40 + i).integers(lo, hi + 1, 1 << 23).astype(np.int16)
                for idx, val in pins:
                    m[idx] = val
                d.set_state(m, 777, 3)
            with Checker(ring) as o:
                o.init(0)
                if m is not None:
                    o.set_state(m, 777, 3)
                r = o.update_blk(sy, n)
                first = 3 if m is not None else 0
                want.append((r, o.get_metrics(), [crc(o.get_row((first + k) % ring)) for k in range(n)], o.min_metric(), o.max_metric()))
        ren = v224.Viterbi224.update_multi_dev(decs, dptr, n)
        for i, d in enumerate(decs):
            first = 3 if states[i] else 0
            assert ren[i] == want[i][0], i
            assert np.array_equal(d.get_metrics(), want[i][1]), i
            assert [crc(d.get_row((first + k) % ring)) for k in range(n)] == want[i][2], i
            assert (d.min_metric(), d.max_metric()) == (want[i][3], want[i][4]), i
        assert decs[1].stats()["sat_stages"] > 0
    finally:
        for d in decs:
            d.delete()
