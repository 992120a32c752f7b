"""Time-segmented decoding of one long soft-symbol stream on several GPUs.

The trellis recursion is serial in time, but survivors forget their start: a decoder started
W stages early from uniform metrics makes the same decisions as the full-stream decoder once
all survivors share an ancestor inside the warm-up.  So a stream of N
bits is cut into G contiguous output ranges; rank g decodes stages [a_g - W, b_g) and emits
the streaming outputs (decodebit(delay, 0) after every stage, vdecode.c:145-152) of [a_g, b_g).
No data-path collective: symbols are scattered by range, output bits gathered by range.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Segment:
    rank: int
    out_first: int      # first output bit index this rank emits
    out_last: int       # one past the last
    stage_first: int    # first trellis stage this rank runs (out_first - warm-up, clamped at 0)
    skip: int           # outputs of the warm-up to discard

    @property
    def nstages(self):
        return self.out_last - self.stage_first


def plan(nbits, world_size, warmup, delay):
    """Split [0, nbits) into world_size contiguous output ranges with a leading warm-up of
    max(warmup, delay) stages (the walk of the first emitted bit must stay inside the segment)."""
    w = max(int(warmup), int(delay))
    base, rem = divmod(int(nbits), int(world_size))
    segs = []
    a = 0
    for g in range(world_size):
        b = a + base + (1 if g < rem else 0)
        s = max(0, a - w)
        segs.append(Segment(g, a, b, s, a - s))
        a = b
    return segs


def decode_segment(decoder, soft, seg, delay):
    """Run one segment on `decoder` (a Viterbi224 with len > delay).  `soft` is the whole stream's
    symbol array (2 per bit) or any array whose index 2*stage addresses the stage's first symbol.
    Returns the uint8 bit array for outputs [seg.out_first, seg.out_last)."""
    if seg.stage_first == 0:
        decoder.init(0)                         # the stream really starts in state 0
    else:
        decoder.init_uniform(5000, -1)          # mid-stream: favour no state
    syms = np.ascontiguousarray(soft[2 * seg.stage_first: 2 * seg.out_last])
    bits, _ = decoder.stream_decode(syms, delay)
    return bits[seg.skip:]


def stitch(parts):
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def decode_distributed(decoder, soft, nbits, delay, warmup, rank, world_size, dist=None):
    """One rank's part of a time-segmented decode plus the gather of all ranks' output bits.
    `dist` is torch.distributed (already initialised; nccl or gloo) or None for world_size 1.
    Returns the full uint8[nbits] output on every rank.  The only communication is the final
    all_gather of decoded bits (1 byte per bit here; nothing crosses ranks inside the hot loop)."""
    segs = plan(nbits, world_size, warmup, delay)
    mine = decode_segment(decoder, soft, segs[rank], delay)
    if dist is None or world_size == 1:
        return mine
    import torch
    longest = max(s.out_last - s.out_first for s in segs)
    buf = torch.zeros(longest, dtype=torch.uint8)
    buf[: mine.size] = torch.from_numpy(mine)
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = buf.to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(outs, buf)
    parts = [o.cpu().numpy()[: s.out_last - s.out_first] for o, s in zip(outs, segs)]
    return stitch(parts)
