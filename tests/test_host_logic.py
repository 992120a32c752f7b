"""CPU tier: host-side logic that sits around the CUDA path -- the vdecode.c symbol pairing / phase
flip mirror, the time-segment planner, and the two-rank (gloo) gather."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import isee3_decoder_b200 as v224

S = v224.streams
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def vdecode_pairs_literal(soft, start_phase=0, dontflip=False):
    """vdecode.c:101-187 transcribed symbol by symbol (test-side restatement; slow on purpose)."""
    sync_vector = S.sync_vector().tolist()
    old = [0] * 4096
    for i in range(0, 4096, 2):
        old[i] = 255 if S.G1FLIP else 0
        old[i + 1] = 255 if S.G2FLIP else 0
    symbols = start_phase
    vdsyms = [0, 0]
    sync_count = 0
    peak_in = peak_out = -1000000
    pairs = []
    for c in soft.tolist():
        old[symbols] = c
        vdsyms[symbols % 2] = c
        if not dontflip:
            ssum = 0
            for k in range(34):
                v = old[(4096 + symbols + k - 33) % 4096] - 128
                ssum += v if sync_vector[k] else -v
            if symbols % 2 == 0:
                peak_out = max(peak_out, ssum)
            else:
                peak_in = max(peak_in, ssum)
                sync_count += 1
                if sync_count >= 2048:
                    sync_count = 0
                    if peak_out > peak_in:
                        symbols = symbols + 1 if symbols % 2 == 0 else symbols - 1
                    peak_in = peak_out = -1000000
        if symbols % 2 == 1:
            pairs.append((vdsyms[0], vdsyms[1]))
        symbols = (symbols + 1) % 4096
    return np.array(pairs, dtype=np.uint8).reshape(-1, 2)


@pytest.mark.parametrize("junk,start_phase,dontflip", [(0, 0, False), (101, 0, False), (7, 1, False), (101, 0, True), (4097, 0, False)])
def test_pair_symbols_equals_literal_vdecode_loop(junk, start_phase, dontflip):
    _, soft = S.telemetry_stream(5 * 1024, 5.0, seed=100 + junk, junk_symbols=junk)
    a = v224.vdecode.pair_symbols(soft, start_phase, dontflip)
    b = vdecode_pairs_literal(soft, start_phase, dontflip)
    assert a.shape == b.shape
    assert np.array_equal(a, b)


def test_phase_flip_happens_once_for_odd_junk():
    _, soft = S.telemetry_stream(6 * 1024, 5.0, seed=5, junk_symbols=33)
    pairs, flips = v224.vdecode.pair_symbols(soft, return_flips=True)
    assert flips == [4095]
    assert pairs.shape[0] == (soft.size - 1) // 2       # one symbol is dropped by the flip
    _, soft = S.telemetry_stream(6 * 1024, 5.0, seed=5, junk_symbols=32)
    _, flips = v224.vdecode.pair_symbols(soft, return_flips=True)
    assert flips == []


def test_segment_plan_covers_stream_exactly_once():
    for nbits, world, warm, delay in [(1000, 1, 64, 24), (1000, 3, 64, 200), (1 << 20, 8, 2048, 200), (17, 4, 8, 4)]:
        segs = v224.segments.plan(nbits, world, warm, delay)
        assert segs[0].out_first == 0 and segs[-1].out_last == nbits
        for a, b in zip(segs, segs[1:]):
            assert a.out_last == b.out_first
        for s in segs:
            assert s.stage_first == max(0, s.out_first - max(warm, delay))
            assert s.skip == s.out_first - s.stage_first
            assert s.nstages == s.out_last - s.stage_first


def test_reencode_tally_is_zero_on_clean_stream():
    bits, soft = S.telemetry_stream(2048, 30.0, seed=9)          # effectively noiseless
    pairs = v224.vdecode.pair_symbols(soft, dontflip=True)
    delay = 64
    lag = delay + 22
    out = np.concatenate([np.zeros(lag, np.uint8), bits])[: pairs.shape[0]]   # what the decoder would emit
    # bit i of `out` after the startup is data bit i - lag; re-encoding it must reproduce the received hard symbols
    errs = v224.vdecode.reencode_symbol_errors(out[delay:], pairs, delay)   # vdecode prints from pair `delay` on
    assert errs == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("conv,expect", [(160, "verified=1 redone=0 spread=0"), (0, "verified=0 redone=1")])
def test_two_rank_verified_segments_gloo(built, conv, expect):
    """world_size 2 over gloo on CPU: segments.decode_verified -- ranges, snapshot exchange, hand-over check, and (conv = 0:
    the check must fail) the exact redo of a range by the previous rank.  The CPU oracle stands in for the GPU decoder
    so that the host protocol of the N>1 path is covered without a GPU; the stitched output equals one sequential decode."""
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_segment_worker.py"), str(conv)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "RESULT diff=0 data_ok=True n=400" in outs[0] and expect in outs[0], outs[0]


def test_library_pairing_equals_the_python_mirror(built):
    """v224x_pair_symbols (the library's host-side entry, run-at-a-time correlation) against the Python mirror of
    vdecode.c:101-140: pairs, comparison symbols of the re-encode tally, flip positions; both start phases, -F, streams
    with junk prefixes and lost symbols at 3 .. -2 dB (false-alarm flips)."""
    for seed, junk, ebn0, nb, drop in [(1, 101, 3.0, 40, []), (2, 0, 1.0, 24, []), (3, 7, 0.0, 20, [9_000]), (4, 1, -2.0, 48, [50_000, 90_001])]:
        _, soft = S.telemetry_stream(nb * 1024, ebn0, seed=seed, junk_symbols=junk)
        soft = np.delete(soft, drop)
        for phase in (0, 1):
            want, wflips = v224.vdecode.pair_symbols(soft, start_phase=phase, return_flips=True)
            got, gflips, cmp_ = v224.pair_symbols(soft, start_phase=phase, delay=200, want_cmp=True)
            fast, fflips = v224.pair_symbols(soft, start_phase=phase)
            assert np.array_equal(got, want) and np.array_equal(fast, want) and gflips == fflips and len(gflips) == len(wflips)
            # comparison symbols: the hard-sliced history 2 * (delay + 22) symbols back, i.e. the pair 222 pairs earlier
            # (exact only while no flip lies in between; checked on the flip-free case)
            if not gflips and phase == 0:
                k = np.arange(300, got.shape[0])
                assert np.array_equal(cmp_[k], (got[k - 222] > 128).astype(np.uint8))
        want = v224.vdecode.pair_symbols(soft, dontflip=True)
        got, gflips = v224.pair_symbols(soft, dontflip=True)
        assert np.array_equal(got, want) and gflips == []
    assert v224.pair_symbols(np.zeros(0, np.uint8))[0].shape == (0, 2)
    assert v224.pair_symbols(np.array([7], np.uint8))[0].shape == (0, 2)


def test_block_driver_pairing_equals_the_python_mirror(built):
    """vdecode_block -P (pairs only, no GPU): the C++ host program's symbol pairing and phase-flip logic (its own
    implementation of vdecode.c:101-140,186) against the Python mirror, which the GPU tier pins to the unmodified
    reference's output: junk prefix (flip), a dropped symbol mid-stream (flip), -p start phase, -F no flipping."""
    import subprocess
    exe = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    bits, soft = v224.streams.telemetry_stream(14 * 1024, 3.0, seed=21, junk_symbols=77)
    soft = np.delete(soft, 15_001)
    for args, kw in ((["-q"], {}), (["-q", "-p"], {"start_phase": 1}), (["-q", "-F"], {"dontflip": True}), (["-q", "-B", "2000"], {})):
        r = subprocess.run([exe, "-P"] + args, input=soft.tobytes(), capture_output=True, timeout=120)
        assert r.returncode == 0, r.stderr
        got = np.frombuffer(r.stdout, dtype=np.uint8).reshape(-1, 2)
        want, flips = v224.vdecode.pair_symbols(soft, return_flips=True, **kw)
        if kw.get("start_phase"):
            want = want.copy(); got = got.copy()
            want[0, 0] = got[0, 0] = 0          # -p: the first pair's even symbol is uninitialised in the reference
        assert got.shape == want.shape and np.array_equal(got, want), args
    r = subprocess.run([exe, "-P"], input=soft.tobytes(), capture_output=True, timeout=120)
    assert r.stderr.count(b"flipping phase") == len(v224.vdecode.pair_symbols(soft, return_flips=True)[1]) >= 2


# ---------------------------------------------------------------------------------------------
# the "next" rows of the scope table: framer and the frame decoder's sync search (host side, no GPU)
# ---------------------------------------------------------------------------------------------
def _host_golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", "host", name + ".npz"))


def test_framer_mode_prints_what_the_reference_framer_prints(built):
    """vdecode_block -f -b (framing only) against the recorded output of the unmodified framer.c (tools/make_golden_host.py):
    junk prefix, five good frames, one frame whose sync word carries a bit error (not reported)."""
    fx = _host_golden("framer_seed7")
    bits = np.unpackbits(fx["bits"])[: int(fx["nbits"])]
    txt = bytes(np.where(bits == 1, ord("1"), ord("0")).astype(np.uint8))
    exe = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    r = subprocess.run([exe, "-f", "-b", "-r", "512"], input=txt, capture_output=True, timeout=60, env=dict(os.environ, LANG="C"))
    assert r.returncode == 0, r.stderr
    assert r.stdout == bytes(fx["stdout"]) and r.stdout.count(b"Frame ") == 5


def frame_sync_literal(soft):
    """decode.c:152-192,270-282 with lock never asserted: per frame, first arg-max of the 34-tap correlator over 2048
    positions, then skip sync_start + 2048 symbols (test-side restatement)."""
    taps = np.where(S.sync_vector() == 1, 1, -1).astype(np.int64)
    x = soft.astype(np.int64) - 128
    base, out = 0, []
    while base + 2048 + 34 <= x.size:
        corr = np.correlate(x[base: base + 2048 + 33], taps, mode="valid")      # corr[i] = sum_k x[base+i+k] * taps[k]
        ss = int(np.argmax(corr[:2048]))
        if base + ss + 2048 + 34 > x.size:
            break
        out.append(base + ss + 34)
        base += ss + 2048
    return out


def test_frame_decoder_sync_search_equals_the_literal_loop(built):
    """decode_block -S (sync search only, no GPU) on the stream of the decode golden fixture and on a noisier one; the
    positions of the good frames are the ones the unmodified decode.c printed."""
    exe = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
    fx = _host_golden("decode_V_seed21")
    _, noisy = S.telemetry_stream(12 * 1024, 1.0, seed=33, junk_symbols=4001)
    for soft in (fx["symbols"], noisy):
        r = subprocess.run([exe, "-S"], input=soft.tobytes(), capture_output=True, timeout=60)
        assert r.returncode == 0, r.stderr
        got = [int(x) for x in r.stdout.split()]
        assert got == frame_sync_literal(soft) and len(got) >= 6
    # the reference's own printout: every frame it found by searching (the first, and each one after a bad frame) is in the list
    ref_lines = [ln for ln in bytes(fx["stdout"]).decode().splitlines() if ln.startswith("Frame ")]
    r = subprocess.run([exe, "-S"], input=fx["symbols"].tobytes(), capture_output=True, timeout=60)
    found = set(int(x) for x in r.stdout.split())
    assert int(ref_lines[0].split()[4]) in found


def test_config5_stream_generator_matches_the_reference_encoder_mirror():
    """tools/config5.py builds its stream with torch ops (on the GPU in the real run): its encoder must be the encoder of
    encode.c:17-35 (numpy mirror in streams.py, itself pinned by the golden fixtures), including the carried-over history,
    and its quantiser must produce symdemod-format bytes."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config5
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, 5000, dtype=np.uint8)
    state = int(rng.integers(0, 1 << 23))
    hist = np.array([(state >> i) & 1 for i in range(22, -1, -1)], dtype=np.uint8)          # oldest first
    s1, s2 = config5.encode(torch, torch.from_numpy(hist), torch.from_numpy(bits))
    want, _ = S.encode_bits(bits, state)
    assert np.array_equal(s1.numpy(), want[0::2]) and np.array_equal(s2.numpy(), want[1::2])
    soft = config5.soften(torch, s1, s2, 60.0, 1, 0).numpy()                                  # noiseless: 128 +- amplitude, clipped
    a, _ = S.symdemod_amplitudes(60.0)
    lo, hi = int(np.clip(128.0 - a, 0, 255)), int(np.clip(128.0 + a, 0, 255))
    assert set(np.unique(soft).tolist()) <= {lo - 1, lo, lo + 1, hi - 1, hi, hi + 1} and np.array_equal(soft > 128, want.astype(bool))


# ---------------------------------------------------------------------------------------------
# sequential decoder of the Fano-first frame policy (host/fano_seq.h) and the time format (host/hostfmt.h)
# ---------------------------------------------------------------------------------------------
def _host_shim():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libhost_shim.so"))
    lib.shim_fano.restype = ctypes.c_int
    lib.shim_fano.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint, ctypes.c_void_p, ctypes.c_int, ctypes.c_ulong, ctypes.c_ulonglong, ctypes.c_ulonglong]
    lib.shim_fano_metric_table.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
    lib.shim_format_hms.argtypes = [ctypes.c_double, ctypes.c_char_p, ctypes.c_int]
    return lib


def test_fano_decoder_metric_tables_and_time_format_equal_the_reference(built):
    """host/fano_seq.h and host/hostfmt.h against values recorded from the unmodified fano.c / metrics.c / timeformat.c
    (tools/make_golden_host.py): metric tables entry for entry; per frame the number of decoded bits, final metric, cycle
    count and decoded bytes -- clean decodes, long searches and timeouts; time stamps across the minute/hour/day carries.
    Where oracle/_ref/libv224_reffano.so is present the same comparison runs live on fresh random frames."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden_host as mg
    fx = _host_golden("fano_cases")
    shim = _host_shim()
    tables = []
    for (sig, noise, bias, scale), want in zip(mg.FANO_TABLES, fx["tables"]):
        t = np.zeros((2, 256), np.int32)
        shim.shim_fano_metric_table(t.ctypes.data, sig, noise, bias, scale)
        assert np.array_equal(t, want), (sig, noise, bias, scale)
        tables.append(t)
    timeouts = 0
    for (name, syms, nbits, ti, delta, maxc, start, tail), want, wdata in zip(mg.fano_cases(), fx["results"], fx["data"]):
        r, metric, cycles, data = mg.run_fano(shim.shim_fano, tables[ti], syms, nbits, delta, maxc, start, tail)
        assert (r, metric, cycles) == tuple(int(x) for x in want), name
        assert np.array_equal(data[: r // 8], wdata[: r // 8]), name
        timeouts += r != nbits
    assert timeouts >= 3
    buf = ctypes.create_string_buffer(64)
    for t, want in zip(fx["times"], fx["hms"]):
        shim.shim_format_hms(float(t), buf, 64)
        assert buf.value.decode() == str(want), t
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libv224_reffano.so")):
        ref = mg.ref_fano_lib()
        rng = np.random.default_rng(77)
        for i in range(40):
            ebn0 = float(rng.uniform(0.5, 4.0))
            bits = S.telemetry_bits(1, rng)
            start = int(rng.integers(0, 1 << 24))
            sym01, _ = S.encode_bits(bits, start)
            syms = S.awgn_symdemod(sym01, ebn0, rng)
            tail = int("".join(map(str, bits[-23:])), 2)
            delta, maxc = int(rng.choice([8, 32, 50])), int(rng.choice([5, 100, 300]))
            a = mg.run_fano(ref.fano, tables[0], syms, 1024, delta, maxc, start, tail)
            b = mg.run_fano(shim.shim_fano, tables[0], syms, 1024, delta, maxc, start, tail)
            assert a[:3] == b[:3] and np.array_equal(a[3][: a[0] // 8], b[3][: b[0] // 8]), (i, ebn0, a[:3], b[:3])


def test_frame_decoder_fano_only_mode_prints_what_the_reference_prints(built):
    """decode_block -F (sync search, lock logic, Fano, printout -- no GPU) against the unmodified `decode -F` where
    oracle/_ref travelled, on streams from 1 to 4 dB: at the low end most frames time out and are printed as partial,
    bad frames; options -n, -p, -r, -s, -d."""
    ref = os.path.join(ROOT, "oracle", "_ref", "decode_sse")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/decode_sse not built (reference checkout absent at build time)")
    exe = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
    env = dict(os.environ, LANG="C")

    def strip(b):
        return [ln.split(b": ", 1)[1] if (b": Fano" in ln or b": Not displaying" in ln) else ln for ln in b.split(b"\n")]
    bad = good = 0
    for ebn0, flags in ((1.0, ["-F"]), (1.5, ["-F", "-p", "-r", "512"]), (2.0, ["-F", "-n"]), (4.0, ["-F", "-s", "4", "-d", "16"])):
        _, soft = S.telemetry_stream(40 * 1024, ebn0, seed=int(ebn0 * 10), junk_symbols=999)
        a = subprocess.run([ref] + flags, input=soft.tobytes(), capture_output=True, env=env, timeout=120)
        b = subprocess.run([exe] + flags + ["-B", "7"], input=soft.tobytes(), capture_output=True, env=env, timeout=120)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert strip(a.stdout) == strip(b.stdout), (ebn0, flags)
        bad += a.stdout.count(b"(bad)")
        good += a.stdout.count(b"Frame ") - a.stdout.count(b"(bad)")
    assert bad >= 20 and good >= 60, (bad, good)


def test_host_programs_on_empty_short_and_ragged_input(built):
    """Edge cases of the host programs that need no GPU: empty input, less than one frame, a stream that ends inside a
    frame, sync words back to back and overlapping the stream start; framer mode against the unmodified framer where
    oracle/_ref travelled."""
    vb = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    db = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
    env = dict(os.environ, LANG="C")
    for cmd in ([vb, "-f", "-b"], [db, "-S"], [vb, "-P", "-q"]):
        r = subprocess.run(cmd, input=b"", capture_output=True, timeout=60, env=env)
        assert r.returncode == 0 and r.stdout == b"", cmd
    r = subprocess.run([db, "-F"], input=b"", capture_output=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.count(b"\n") == 2 and b"Frame" not in r.stdout          # the two banner lines only
    _, soft = S.telemetry_stream(3 * 1024, 5.0, seed=8)
    for n in (1, 2081, 2082, 4000, soft.size - 1):                       # 2082 = one frame + sync: the least the search needs (decode.c:152-161)
        r = subprocess.run([db, "-S"], input=soft[:n].tobytes(), capture_output=True, timeout=60, env=env)
        assert r.returncode == 0 and [int(x) for x in r.stdout.split()] == frame_sync_literal(soft[:n]), n
    ref = os.path.join(ROOT, "oracle", "_ref", "framer_ref")
    sync = [(S.SYNCWORD >> (39 - i)) & 1 for i in range(40)]
    rng = np.random.default_rng(5)
    cases = [sync, sync + sync, [0] * 7 + sync + [1] + sync, sync[1:] + sync, list(rng.integers(0, 2, 1500)) + sync + list(rng.integers(0, 2, 1024 - 40)) + sync]
    for bits in cases:
        txt = bytes(ord("1") if b else ord("0") for b in bits)
        got = subprocess.run([vb, "-f", "-b"], input=txt, capture_output=True, timeout=60, env=env)
        assert got.returncode == 0
        nfr = got.stdout.count(b"Frame ")
        assert nfr >= 1 and got.stdout.count(b"\n") == nfr * 10          # header + 8 hex lines + blank line per frame
        if os.path.exists(ref):
            want = subprocess.run([ref], input=txt, capture_output=True, timeout=60, env=env)
            assert got.stdout == want.stdout, bits[:50]


def _feed_slowly(cmd, data, chunks, env=None):
    """Run cmd with `data` trickling into its stdin in pieces (a live pipe: reads return short counts, polls find nothing waiting)."""
    import time
    p = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    import threading
    out = {}
    t = threading.Thread(target=lambda: out.update(stdout=p.stdout.read(), stderr=p.stderr.read()))
    t.start()
    pos = 0
    for n in chunks:
        if pos >= len(data):
            break
        p.stdin.write(data[pos: pos + n]); p.stdin.flush()
        pos += n
        time.sleep(0.001)
    p.stdin.write(data[pos:])
    p.stdin.close()
    t.join(120)
    assert p.wait(60) == 0, out.get("stderr")
    return out["stdout"]


def test_host_programs_give_the_same_output_on_a_trickling_pipe(built):
    """The block drivers read with read(2) and extend a speculative frame batch only with frames that have already
    arrived; what they print must not depend on how the input is chopped up in time."""
    vb = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    db = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
    env = dict(os.environ, LANG="C")
    rng = np.random.default_rng(12)
    _, soft = S.telemetry_stream(30 * 1024, 2.0, seed=41, junk_symbols=321)
    data = soft.tobytes()
    chunks = [int(x) for x in rng.integers(1, 5000, 400)]
    for cmd in ([db, "-F"], [db, "-F", "-n", "-B", "4"], [db, "-S"], [vb, "-P", "-q"], [vb, "-P", "-q", "-B", "1500"]):
        whole = subprocess.run(cmd, input=data, capture_output=True, timeout=120, env=env)
        assert whole.returncode == 0
        assert _feed_slowly(cmd, data, chunks, env) == whole.stdout, cmd
    bits = np.unpackbits(_host_golden("framer_seed7")["bits"])
    txt = bytes(np.where(bits == 1, ord("1"), ord("0")).astype(np.uint8))
    whole = subprocess.run([vb, "-f", "-b"], input=txt, capture_output=True, timeout=60, env=env)
    assert _feed_slowly([vb, "-f", "-b"], txt, [int(x) for x in rng.integers(1, 300, 100)], env) == whole.stdout
