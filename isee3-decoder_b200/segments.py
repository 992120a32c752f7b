"""Time-segmented decoding of one long pair stream by several ranks (one process per GPU, torch.distributed).

The trellis recursion is serial in time, but survivors forget their start: a decoder started W = delay + conv stages
early from uniform metrics makes the decisions of the full-stream decoder from the stage at which the two path-metric
vectors differ by a constant.  A stream of N pairs is cut into G contiguous output ranges; rank g
decodes stages [a_g - W, b_g) and emits the streaming outputs (decodebit(delay, 0) after every stage, vdecode.c:145-152)
of [a_g, b_g).  The hand-over between two ranks is VERIFIED, the way v224x_stream_decode_seg verifies the hand-overs
between its lockstep decoders on one GPU and v224x_multi_stream_decode those between the GPUs of one process:

  * rank g saves its metrics at stream position a_g - delay (`delay` stages before its first output: every decision row
    its tracebacks touch lies after that point), rank g-1 saves its metrics at the same position;
  * rank g sends the 16 MiB snapshot to rank g-1 (the one exchange step of the path: NCCL send/recv over NVLink, once per
    range), which checks that the two vectors differ by a constant (v224x_metric_spread_dev == 0);
  * if they do not, the rank that holds the exact state at a_g decodes range g again from there (lead = 0), and its
    output replaces rank g's.  The stitched output is therefore what ONE sequential decoder produces: residual
    differences are 0 by construction, and the number of verified / redone hand-overs is reported.

`decode_verified` is written against a small decoder interface so that the CPU test tier can run the whole protocol,
including the redo path, over gloo with the CPU oracle standing in for the GPU (tests/mp_segment_worker.py).
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
    segs = []
    for g in range(world_size):
        a = int(nbits) * g // int(world_size)
        b = int(nbits) * (g + 1) // int(world_size)
        s = max(0, a - w)
        segs.append(Segment(g, a, b, s, a - s))
    return segs


def decode_segment(decoder, soft, seg, delay):
    """Run one segment on `decoder` (a Viterbi224 with len > delay), unverified.  `soft` is the whole stream's
    symbol array (2 per bit).  Returns the uint8 bit array for outputs [seg.out_first, seg.out_last)."""
    if seg.stage_first == 0:
        decoder.init(0)                         # the stream really starts in state 0
    else:
        decoder.init_uniform(5000, -1)          # mid-stream: favour no state
    syms = np.ascontiguousarray(soft[2 * seg.stage_first: 2 * seg.out_last])
    bits, _ = decoder.stream_decode(syms, delay)
    return bits[seg.skip:]


def stitch(parts):
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


class GpuRangeDecoder:
    """The decoder interface of decode_verified over a Viterbi224 handle: ranges through v224x_range_decode_dev,
    snapshots in torch device tensors (so that torch.distributed can move them), spread through v224x_metric_spread_dev."""

    def __init__(self, dec, torch, device, nseg=3):
        self.dec, self.torch, self.device, self.nseg = dec, torch, device, nseg
        nb = int(dec.lib.v224x_snapshot_bytes())
        self.snap_early = torch.empty(nb, dtype=torch.uint8, device=device)
        self.snap_late = torch.empty(nb, dtype=torch.uint8, device=device)
        self.snap_peer = torch.empty(nb, dtype=torch.uint8, device=device)
        self.reports = []

    def start_of_stream(self):
        self.dec.init(0)

    def range_decode(self, syms_t, lead, nout, delay, conv, bits_t, want_early, want_late):
        """syms_t / bits_t: uint8 torch tensors on this GPU (2 * (lead + nout) symbols in, nout bits out)."""
        rep = self.dec.range_decode_dev(syms_t.data_ptr(), lead, nout, delay, bits_t.data_ptr(), self.nseg, conv,
                                        self.snap_early.data_ptr() if want_early else None,
                                        self.snap_late.data_ptr() if want_late else None)
        self.reports.append(rep)
        return rep

    def spread(self, late, other):
        self.torch.cuda.synchronize(self.device)             # the snapshot arrived on torch's stream
        return self.dec.metric_spread_dev(late.data_ptr(), other.data_ptr())


def decode_verified(rd, load_range, nbits, delay, conv, rank, world, dist, torch, ctrl_device="cpu"):
    """One rank's part of a verified time-segmented decode.

    rd          : range decoder (GpuRangeDecoder, or a CPU stand-in with the same methods)
    load_range  : load_range(stage_first, stage_last) -> (symbols, bits): uint8 torch tensors where rd wants them (the
                  rank's GPU; CPU for the stand-in): the symbols of stream stages [stage_first, stage_last) and room for
                  stage_last - stage_first output bits
    Returns (first, bits, report): this rank's output range starts at stream bit `first`; bits[:n] are its exact bits
    (its own decode, or the previous owner's re-decode received over dist); report counts the hand-overs."""
    W = delay + conv
    segs = plan(nbits, world, W, delay)
    me = segs[rank]
    lead = me.out_first - me.stage_first
    if world > 1 and rank > 0 and lead != W:
        raise ValueError("stream too short for this many ranks: a range must start at least delay + conv stages into the stream")
    nout = me.out_last - me.out_first
    syms_h, bits_h = load_range(me.stage_first, me.out_last)
    if rank == 0:
        rd.start_of_stream()
    rd.range_decode(syms_h, lead, nout, delay, conv, bits_h, want_early=rank > 0, want_late=rank + 1 < world)
    report = {"ranks": world, "handovers_verified": 0, "ranges_redone": 0, "worst_spread": 0, "residual_diffs": 0}
    if world == 1:
        return me.out_first, bits_h, report

    def as_int(t):
        return int(t.item())

    # ---- first round: every hand-over in parallel (rank g -> rank g-1) ----
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, rd.snap_early, rank - 1))
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.irecv, rd.snap_peer, rank + 1))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    mine = torch.zeros(world, dtype=torch.int64, device=ctrl_device)
    if rank + 1 < world:
        mine[rank + 1] = rd.spread(rd.snap_late, rd.snap_peer)           # hand-over into range rank+1
    dist.all_reduce(mine, op=dist.ReduceOp.MAX)
    spreads = [as_int(x) for x in mine]
    # ---- in stream order: who holds the exact output of every range ----
    owner = [0] * world
    for g in range(1, world):
        e = owner[g - 1]
        ok = spreads[g] == 0
        if e != g - 1:
            # range g-1 was decoded again by rank e: rank g's decoder has to agree with THAT decoder
            if rank == g:
                dist.send(rd.snap_early, e)
            flag = torch.zeros(1, dtype=torch.int64, device=ctrl_device)
            if rank == e:
                dist.recv(rd.snap_peer, g)
                flag[0] = rd.spread(rd.snap_late, rd.snap_peer)
            dist.broadcast(flag, e)
            spreads[g] = as_int(flag)
            ok = spreads[g] == 0
        report["worst_spread"] = max(report["worst_spread"], spreads[g])
        if ok:
            report["handovers_verified"] += 1
            owner[g] = g
            continue
        # rank e is exact at the start of range g: it decodes the range again, continuing its state
        report["ranges_redone"] += 1
        owner[g] = e
        sg = segs[g]
        if rank == e:
            s2, b2 = load_range(sg.out_first, sg.out_last)
            rd.range_decode(s2, 0, sg.out_last - sg.out_first, delay, conv, b2, want_early=False, want_late=g + 1 < world)
            dist.send(b2[: sg.out_last - sg.out_first].contiguous(), g)
        if rank == g:
            tmp = torch.empty_like(bits_h[:nout])
            dist.recv(tmp, e)
            bits_h[:nout].copy_(tmp)
    return me.out_first, bits_h, report
