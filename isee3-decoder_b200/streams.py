"""Synthetic decoder input, generated the way the reference's test programs do it.

* encode():           the K=24 r=1/2 encoder of encode.c:17-35 (MSB-first, POLY1 symbol first,
                      second symbol inverted -- code.h:59-63), vectorised with numpy.
* awgn_vtest():       vtest224.c:93-112 / sim.c:17-51 -- BPSK +-Gain around 128 plus Gaussian
                      noise, quantised to 8 bits by the CDF-bin rule of sim.c (bin s covers
                      (s-128-0.5, s-128+0.5]).
* awgn_symdemod():    symdemod.c:190,240-251 wire format -- total RMS amplitude 100, +128, clipped
                      to [0,255] and truncated.
* vtest_frame():      one vtest224 frame (random payload, K zero tail bits), config 1.
* telemetry_stream(): 1024-bit minor frames ending in the 40-bit sync word 0x12fc819fbe
                      (decode.c:21-24), continuous encoder state, config 2.

All randomness comes from numpy's seeded PCG64; parity with the reference is judged on
identical symbol BYTES, never on identical random numbers.
"""
import numpy as np

K = 24
POLY1 = 0o73665667
POLY2 = 0o73665665
G1FLIP = 0
G2FLIP = 1
FRAMEBITS = 1024
SYNCWORD = 0x12FC819FBE
SYNCBITS_IN_FRAME = 40


def bytes_to_bits(data):
    """MSB-first bit expansion (encode.c:26)."""
    return np.unpackbits(np.ascontiguousarray(data, dtype=np.uint8))


def bits_to_bytes(bits):
    return np.packbits(np.ascontiguousarray(bits, dtype=np.uint8))


def encode_bits(bits, encstate=0):
    """Encode a 0/1 bit array.  Returns (symbols uint8[2n] of 0/1, final 24-bit encoder state)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    n = bits.size
    hist = np.array([(encstate >> i) & 1 for i in range(K - 2, -1, -1)], dtype=np.uint8)   # oldest first
    d = np.concatenate([hist, bits])
    # register bit i at time t is d[t - i]  (bit 0 = newest)
    s1 = np.zeros(n, dtype=np.uint8)
    s2 = np.zeros(n, dtype=np.uint8)
    for i in range(K):
        if i > K - 1:
            break
        seg = d[K - 1 - i: K - 1 - i + n]
        if (POLY1 >> i) & 1:
            s1 ^= seg
        if (POLY2 >> i) & 1:
            s2 ^= seg
    s1 ^= G1FLIP
    s2 ^= G2FLIP
    out = np.empty(2 * n, dtype=np.uint8)
    out[0::2] = s1
    out[1::2] = s2
    tail = d[-K:] if d.size >= K else np.concatenate([np.zeros(K - d.size, np.uint8), d])
    state = 0
    for b in tail:
        state = (state << 1) | int(b)
    return out, state & ((1 << K) - 1)


def encode(data_bytes, encstate=0):
    """encode.c:17-35 on a byte array."""
    return encode_bits(bytes_to_bits(data_bytes), encstate)


def vtest_noise_sigma(ebn0_db, gain=24.0, rate=0.5):
    """vtest224.c:93-95."""
    esn0 = ebn0_db + 10 * np.log10(rate)
    return gain * np.sqrt(0.5) / 10 ** (0.05 * esn0)


def awgn_vtest(symbols01, ebn0_db, rng, gain=24.0):
    """sim.c:17-51 in distribution: 8-bit quantised BPSK+AWGN around 128."""
    sigma = vtest_noise_sigma(ebn0_db, gain)
    y = (2.0 * symbols01.astype(np.float64) - 1.0) * gain + sigma * rng.standard_normal(symbols01.size)
    return np.clip(np.ceil(y + 127.5), 0, 255).astype(np.uint8)


def symdemod_amplitudes(ebn0_db, total_rms=100.0, rate=0.5):
    """Signal / noise amplitude with total RMS 100 (symdemod.c:190; decode.c:125-131)."""
    esn0 = 10 ** (ebn0_db / 10.0) * rate
    sigma = total_rms / np.sqrt(1.0 + 2.0 * esn0)
    return sigma * np.sqrt(2.0 * esn0), sigma


def awgn_symdemod(symbols01, ebn0_db, rng):
    """symdemod.c:240-251: scaled = gain*integrator + 128, clipped to [0,255], truncated."""
    a, sigma = symdemod_amplitudes(ebn0_db)
    y = (2.0 * symbols01.astype(np.float64) - 1.0) * a + sigma * rng.standard_normal(symbols01.size) + 128.0
    return np.clip(y, 0, 255).astype(np.uint8)


def vtest_frame(framebits, ebn0_db, seed, gain=24.0):
    """One vtest224 BER-mode frame (vtest224.c:100-112).  Returns (data bytes, soft symbols)."""
    rng = np.random.default_rng(seed)
    nbytes = framebits // 8
    data = np.zeros(nbytes, dtype=np.uint8)
    npay = (framebits - K) // 8
    data[:npay] = rng.integers(0, 256, npay, dtype=np.uint8)
    sym01, _ = encode(data, 0)
    return data, awgn_vtest(sym01, ebn0_db, rng, gain)


def telemetry_bits(nframes, rng):
    """nframes minor frames of 1024 bits, the last 40 bits of each = SYNCWORD (decode.c:21-24,241-246)."""
    bits = rng.integers(0, 2, (nframes, FRAMEBITS), dtype=np.uint8)
    sync = np.array([(SYNCWORD >> (SYNCBITS_IN_FRAME - 1 - i)) & 1 for i in range(SYNCBITS_IN_FRAME)], dtype=np.uint8)
    bits[:, FRAMEBITS - SYNCBITS_IN_FRAME:] = sync
    return bits.reshape(-1)


def telemetry_stream(nbits, ebn0_db, seed, junk_symbols=0, style="symdemod"):
    """Config-2 style stream: framed telemetry, continuous encoder, soft symbols in symdemod
    (or vtest) format, optionally preceded by `junk_symbols` noise-only symbols (an odd count
    forces vdecode's phase flip).  Returns (data bits, soft symbols)."""
    rng = np.random.default_rng(seed)
    nframes = (nbits + FRAMEBITS - 1) // FRAMEBITS
    bits = telemetry_bits(nframes, rng)[:nbits]
    sym01, _ = encode_bits(bits, 0)
    soft = awgn_symdemod(sym01, ebn0_db, rng) if style == "symdemod" else awgn_vtest(sym01, ebn0_db, rng)
    if junk_symbols:
        _, sigma = symdemod_amplitudes(ebn0_db)
        junk = np.clip(128.0 + sigma * rng.standard_normal(junk_symbols), 0, 255).astype(np.uint8)
        soft = np.concatenate([junk, soft])
    return bits, soft


def segment_bits(seed, segment_index, nbits):
    """Data bits of segment `segment_index` of an endless seeded stream (multi-GPU bench): any
    rank can regenerate any segment, so warm-up prefixes need no communication."""
    rng = np.random.default_rng([seed, segment_index])
    return rng.integers(0, 2, nbits, dtype=np.uint8)


def sync_vector():
    """The 34 encoded sync symbols the reference hard-codes (vdecode.c:27-30): the last 34
    symbols of encode(SYNCWORD)."""
    sym, _ = encode(np.array([(SYNCWORD >> (8 * i)) & 0xFF for i in range(4, -1, -1)], dtype=np.uint8), 0)
    return sym[-34:]
