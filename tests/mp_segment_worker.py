"""Worker of tests/test_host_logic.py::test_two_rank_segmented_decode_gloo: one rank of a world_size-2
gloo job.  Decodes its time segment with the CPU oracle standing in for the GPU decoder (the
host-side partition / gather logic is what is under test here), gathers, compares with a single pass."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import isee3_decoder_b200 as v224   # noqa: E402
import pyoracle                     # noqa: E402


class CpuDecoder(pyoracle.Oracle):
    """Adds the one extension the segment code needs (v224x_init_uniform) to the CPU oracle."""

    def init_uniform(self, bias=5000, start_state=-1):
        self.init(0)
        m = np.full(1 << 23, -32768 + bias, dtype=np.int16)
        if start_state >= 0:
            m[start_state] = -32768
        self.set_state(m, 0, 0)
        return 0


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, delay, warm = 144, 24, 72
    bits, soft = v224.streams.telemetry_stream(n, 8.0, seed=3)
    with CpuDecoder(delay + n) as d:
        out = v224.segments.decode_distributed(d, soft, n, delay, warm, rank, world, dist)
    ok = True
    if rank == 0:
        with CpuDecoder(delay + n) as d:
            d.init(0)
            full, _ = d.stream_decode(soft, delay)
        diff = int((out != full).sum())
        lag = delay + 22
        data_ok = bool(np.array_equal(full[lag:], bits[: n - lag]))
        print(f"RESULT diff={diff} data_ok={data_ok} n={out.size}", flush=True)
        ok = diff == 0 and data_ok and out.size == n
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
