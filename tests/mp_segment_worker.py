"""Worker of tests/test_host_logic.py::test_two_rank_verified_segments_gloo: one rank of a world_size-2 gloo job running
isee3-decoder_b200/segments.py::decode_verified -- ranges, snapshot exchange, hand-over check, redo of a failed range --
with the CPU oracle standing in for the GPU decoder (the host-side protocol is what is under test here).
argv[1] = conv (stages a late-started decoder gets before the check; 0 makes the check fail and forces the redo path)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import isee3_decoder_b200 as v224   # noqa: E402
import pyoracle                     # noqa: E402

NS = 1 << 23


class CpuRangeDecoder:
    """segments.GpuRangeDecoder's interface over the CPU oracle: v224x_range_decode semantics stage by stage."""

    def __init__(self, ring):
        self.o = pyoracle.Oracle(ring)
        self.snap_early = torch.zeros(NS, dtype=torch.int16)
        self.snap_late = torch.zeros(NS, dtype=torch.int16)
        self.snap_peer = torch.zeros(NS, dtype=torch.int16)

    def start_of_stream(self):
        self.o.init(0)

    def range_decode(self, syms_t, lead, nout, delay, conv, bits_t, want_early, want_late):
        if lead > 0:                                      # v224x_init_uniform(5000, -1): no state is favoured
            self.o.init(0)
            self.o.set_state(np.full(NS, -32768 + 5000, dtype=np.int16), 0, 0)
        syms = syms_t.numpy()
        for i in range(lead + nout + 1):
            if want_early and i == lead - delay:
                self.snap_early.copy_(torch.from_numpy(self.o.get_metrics()))
            if want_late and i == lead + nout - delay:
                self.snap_late.copy_(torch.from_numpy(self.o.get_metrics()))
            if i == lead + nout:
                break
            self.o.update_blk(syms[2 * i: 2 * i + 2], 1)
            if i >= lead:
                bits_t[i - lead] = self.o.decodebit(delay, 0) & 0xFF

    def spread(self, late, other):
        d = late.to(torch.int32) - other.to(torch.int32)
        return int(d.max() - d.min())


def main():
    conv = int(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, delay = 400, 24
    bits, soft = v224.streams.telemetry_stream(n, 8.0, seed=3)
    soft_t = torch.from_numpy(np.ascontiguousarray(soft))

    def load_range(a, b):
        return soft_t[2 * a: 2 * b].clone(), torch.zeros(b - a, dtype=torch.uint8)

    rd = CpuRangeDecoder(delay + n)
    first, mine, rep = v224.segments.decode_verified(rd, load_range, n, delay, conv, rank, world, dist, torch)
    segs = v224.segments.plan(n, world, delay + conv, delay)
    nout = segs[rank].out_last - segs[rank].out_first
    parts = [torch.zeros(s.out_last - s.out_first, dtype=torch.uint8) for s in segs]
    dist.all_gather(parts, mine[:nout].contiguous())            # equal range lengths here (n divisible by world)
    out = torch.cat(parts).numpy()
    ok = True
    if rank == 0:
        with pyoracle.Oracle(delay + n) as d:
            d.init(0)
            full, _ = d.stream_decode(soft, delay)
        diff = int((out != full).sum())
        lag = delay + 22
        data_ok = bool(np.array_equal(full[lag:], bits[: n - lag]))
        print(f"RESULT diff={diff} data_ok={data_ok} n={out.size} verified={rep['handovers_verified']} redone={rep['ranges_redone']} "
              f"spread={rep['worst_spread']}", flush=True)
        ok = diff == 0 and data_ok and out.size == n
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
