"""Block-mode mirror of the reference's streaming driver vdecode.c (vdecode.c:38-189).

vdecode.c feeds the decoder one symbol pair at a time and asks for one bit back per pair
(update(1) + decodebit(delay, 0), vdecode.c:145-152).  Which symbols form a pair is decided
by its sync correlator (vdecode.c:107-140), and that decision uses received symbols only --
never decoder output -- so the whole pair sequence, including the one garbage pair produced
at every phase flip, can be worked out on the host first and handed to the GPU as one block.
`pair_symbols()` does that; `vdecode()` returns exactly the characters vdecode prints.
"""
import numpy as np

from .streams import sync_vector, K, POLY1, POLY2, G1FLIP, G2FLIP

FRAMESYMBOLS = 2048          # vdecode.c:15
SYMBOLBUFSIZE = 4096         # vdecode.c:18
NSYNC = 34                   # vdecode.c:16,27-30


def pair_symbols(soft, start_phase=0, dontflip=False, return_flips=False):
    """Replay vdecode.c:101-140,186 over the byte stream `soft`.

    Returns the (npairs, 2) uint8 array of symbol pairs given to update_viterbi224_blk, in order.
    start_phase=1 is `-p` (vdecode.c:77), dontflip is `-F` (vdecode.c:71)."""
    soft = np.ascontiguousarray(soft, dtype=np.uint8)
    n = soft.size
    sv = 2.0 * sync_vector().astype(np.float64) - 1.0          # +1 where sync_vector[k] else -1 (vdecode.c:112-116)
    # hist[r + PRE] = content of ring slot r (r counted without the modulo; the window is 34 << 4096)
    PRE = NSYNC
    hist = np.empty(n + PRE + 2, dtype=np.float64)
    # preset ring (vdecode.c:55-58): even slots G1FLIP?255:0, odd slots G2FLIP?255:0
    idx = np.arange(-PRE, n + 2)
    hist[:] = np.where(idx % 2 == 0, 255.0 if G1FLIP else 0.0, 255.0 if G2FLIP else 0.0)
    pairs = []
    flips = []
    r = int(start_phase)          # ring slot of the next input symbol
    pos = 0                       # next input index
    sync_count = 0
    peak_in = peak_out = -1000000
    even_sym = 0                  # vdsyms[0]; (uninitialised in the reference before the first even symbol)
    while pos < n:
        if dontflip:
            count = n - pos
        else:
            odd_needed = FRAMESYMBOLS - sync_count
            count = 2 * odd_needed if r % 2 == 0 else 2 * odd_needed - 1
            count = min(count, n - pos)
        chunk = soft[pos:pos + count]
        hist[r + PRE: r + PRE + count] = chunk
        rs = np.arange(r, r + count)
        odd = rs % 2 == 1
        if not dontflip:
            # sync_sum at slot s = sum_k +-(ring[s+k-33] - 128)
            win = hist[r + PRE - (NSYNC - 1): r + PRE + count] - 128.0
            sums = np.correlate(win, sv, mode="valid")                      # length == count
            if (~odd).any():
                peak_out = max(peak_out, int(round(sums[~odd].max())))
            if odd.any():
                peak_in = max(peak_in, int(round(sums[odd].max())))
            sync_count += int(odd.sum())
        # pairs decoded in this chunk: every odd slot, paired with the even slot before it
        odd_slots = rs[odd]
        evens = np.where(odd_slots - 1 >= r, hist[np.maximum(odd_slots - 1, -PRE) + PRE], float(even_sym))
        # the even symbol preceding the chunk's first odd slot may predate this chunk
        if odd_slots.size and odd_slots[0] - 1 < r:
            evens[0] = even_sym
        p = np.stack([evens, hist[odd_slots + PRE]], axis=1).astype(np.uint8)
        last_even = rs[~odd]
        if last_even.size:
            even_sym = int(hist[last_even[-1] + PRE])
        pos += count
        r += count
        flipped = False
        if not dontflip and sync_count >= FRAMESYMBOLS:
            sync_count = 0
            if peak_out > peak_in:
                # vdecode.c:126-133: the current (odd-slot) symbol is not decoded; the next input
                # overwrites its slot and is decoded with the stale even symbol.
                flipped = True
                flips.append(pos - 1)
                p = p[:-1]
                r -= 1
            peak_in = peak_out = -1000000
        if p.size:
            pairs.append(p)
    out = np.concatenate(pairs, axis=0) if pairs else np.zeros((0, 2), np.uint8)
    return (out, flips) if return_flips else out


def vdecode(decoder_factory, soft, delay=200, start_phase=0, dontflip=False):
    """What `vdecode -d delay [-p] [-F]` writes to stdout for the byte stream `soft`: an ASCII
    '0'/'1' array, one character per decoded bit, the first `delay` bits suppressed
    (vdecode.c:151-158).  `decoder_factory(len)` must return a Viterbi224-like object."""
    if delay < 24:                                  # vdecode.c:86-88
        delay = 200
    pairs = pair_symbols(soft, start_phase, dontflip)
    npairs = pairs.shape[0]
    if npairs == 0:
        return np.zeros(0, dtype=np.uint8)
    block = min(npairs, 8192)
    dec = decoder_factory(delay + block)            # vdecode.c:94 uses delay+1; a block-mode ring needs block more rows
    try:
        dec.init(0)                                 # vdecode.c:96
        bits, _ = dec.stream_decode(pairs.reshape(-1), delay)
    finally:
        dec.delete()
    return (bits[delay:] + ord("0")).astype(np.uint8)


def reencode_symbol_errors(bits_out, pairs, delay):
    """The re-encode tally of vdecode.c:155-177: decoded bit i is re-encoded and compared with the
    hard-sliced symbols received delay+K-2 pairs earlier.  Returns the total symbol error count
    over the bits where both exist (the reference prints it per status interval)."""
    bits = np.ascontiguousarray(bits_out, dtype=np.uint8)
    n = bits.size
    reg_hist = np.concatenate([np.zeros(K - 1, np.uint8), bits])
    s1 = np.zeros(n, np.uint8)
    s2 = np.zeros(n, np.uint8)
    for i in range(K):
        seg = reg_hist[K - 1 - i: K - 1 - i + n]
        if (POLY1 >> i) & 1:
            s1 ^= seg
        if (POLY2 >> i) & 1:
            s2 ^= seg
    s1 ^= G1FLIP
    s2 ^= G2FLIP
    # output bit i was produced after pair (delay + i); it corresponds to pair (delay + i) - (delay + K - 2)
    lag = K - 2
    src = np.arange(n) - lag
    ok = src >= 0
    hard = pairs[src[ok]] > 128
    return int((s1[ok] ^ hard[:, 0]).sum() + (s2[ok] ^ hard[:, 1]).sum())
