"""Shared test vocabulary: a "script" is a list of ABI calls, run unchanged against the GPU
library (through the C ABI), the CPU oracle, or the unmodified reference.  The recorded
results of the reference are the golden fixtures in tests/golden/ (tools/make_golden.py)."""
import json
import zlib

import numpy as np


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def run_script(factory, script, syms, inspect=True):
    """factory(len) -> decoder object with the ABI-mirroring methods.  Returns a JSON-able dict."""
    res = []
    dec = None
    rows_written = 0
    length = 0
    try:
        for op in script:
            name = op[0]
            if name == "create":
                length = op[1]
                dec = factory(length)
            elif name == "init":
                res.append(["init", dec.init(op[1])])
                rows_written = 0
            elif name == "update":
                start, n = op[1], op[2]
                r = dec.update_blk(syms[2 * start: 2 * (start + n)], n)
                rows_written += n
                res.append(["update", int(r)])
            elif name == "chainback":
                out = dec.chainback(op[1], op[2])
                res.append(["chainback", out.tobytes().hex()])
            elif name == "decodebit":
                res.append(["decodebit", int(dec.decodebit(op[1], op[2]))])
            elif name == "decodeword":
                res.append(["decodeword", int(dec.decodeword(op[1], op[2]))])
            elif name == "minmax":
                res.append(["minmax", int(dec.min_metric()), int(dec.max_metric())])
            elif name == "set_state":
                # deterministic synthetic mid-stream state: metrics uniform in [lo, hi], a few pinned entries
                seed, lo, hi, pins, renormals, stages = op[1:7]
                m = np.random.default_rng(seed).integers(lo, hi + 1, 1 << 23).astype(np.int16)
                for idx, val in pins:
                    m[idx] = val
                dec.set_state(m, renormals, stages)
                rows_written = max(rows_written, stages)
                res.append(["set_state", crc(m)])
            elif name == "stream":
                start, n, delay = op[1], op[2], op[3]
                bits, r = dec.stream_decode(syms[2 * start: 2 * (start + n)], delay, n)
                rows_written += n
                res.append(["stream", bits.tobytes().hex(), int(r)])
            else:
                raise ValueError(name)
        final = {}
        if inspect:
            m = dec.get_metrics()
            final["metrics_crc"] = crc(m)
            final["metrics_min"] = int(m.min())
            final["metrics_max"] = int(m.max())
            nrows = min(rows_written, length)
            final["row_crcs"] = [crc(dec.get_row(r)) for r in range(nrows)]
        return {"results": res, "final": final}
    finally:
        if dec is not None:
            dec.delete()


def save_case(path, name, script, syms, outcome, meta):
    np.savez_compressed(path, name=name, script=json.dumps(script), syms=np.ascontiguousarray(syms, dtype=np.uint8),
                        outcome=json.dumps(outcome), meta=json.dumps(meta))


def load_case(path):
    z = np.load(path, allow_pickle=False)
    return {"name": str(z["name"]), "script": json.loads(str(z["script"])), "syms": z["syms"],
            "outcome": json.loads(str(z["outcome"])), "meta": json.loads(str(z["meta"]))}


def compare_outcomes(got, want, what=""):
    """Assert equality with a readable first difference."""
    assert len(got["results"]) == len(want["results"]), f"{what}: result count"
    for i, (g, w) in enumerate(zip(got["results"], want["results"])):
        assert g == w, f"{what}: op #{i} {w[0]}: got {str(g)[:120]} want {str(w)[:120]}"
    if want["final"] and got["final"]:
        for k in ("metrics_min", "metrics_max", "metrics_crc"):
            assert got["final"][k] == want["final"][k], f"{what}: final {k}: got {got['final'][k]} want {want['final'][k]}"
        gr, wr = got["final"]["row_crcs"], want["final"]["row_crcs"]
        assert len(gr) == len(wr), f"{what}: row count"
        bad = [i for i, (a, b) in enumerate(zip(gr, wr)) if a != b]
        assert not bad, f"{what}: {len(bad)} decision rows differ, first at row {bad[0]}"
