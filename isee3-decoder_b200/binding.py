"""ctypes binding of libviterbi224_b200.so.

Method names and argument meaning follow the reference's C ABI (viterbi224.h:8-16) so that
tests read like the reference's own callers (vtest224.c:116-118, vdecode.c:94-152)."""
import ctypes
import os

import numpy as np

NSTATES = 1 << 23
ROWWORDS = 1 << 18

ABI_SYMBOLS = ["create_viterbi224", "init_viterbi224", "update_viterbi224_blk", "chainback_viterbi224",
               "decodebit_viterbi224", "decodeword_viterbi224", "max_metric_viterbi224", "min_metric_viterbi224",
               "delete_viterbi224"]
EXT_SYMBOLS = ["v224x_device_count", "v224x_set_device", "v224x_last_error", "v224x_version", "v224x_stream_decode",
               "v224x_stream_decode_dev", "v224x_stream_decode_seg", "v224x_stream_decode_seg_dev", "v224x_decode_frames", "v224x_update_dev", "v224x_update_multi_dev", "v224x_init_uniform", "v224x_dev_alloc", "v224x_dev_free",
               "v224x_h2d", "v224x_d2h", "v224x_host_alloc_pinned", "v224x_host_free_pinned", "v224x_timer_start",
               "v224x_timer_stop_ms", "v224x_kernel_time_reset", "v224x_kernel_time_enable", "v224x_kernel_time_ms", "v224x_kernel_time_passes",
               "v224x_get_stats", "v224x_get_metrics", "v224x_set_state", "v224x_get_row", "v224x_set_option",
               "v224x_pair_symbols", "v224x_range_decode", "v224x_range_decode_dev", "v224x_metric_spread_dev", "v224x_snapshot_bytes",
               "v224x_multi_create", "v224x_multi_init", "v224x_multi_stream_decode", "v224x_multi_delete", "v224x_trim"]


class V224Error(RuntimeError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_ulonglong), ("fused_passes", ctypes.c_ulonglong), ("careful_passes", ctypes.c_ulonglong),
                ("single_stages", ctypes.c_ulonglong), ("sat_stages", ctypes.c_ulonglong), ("invalidated_passes", ctypes.c_ulonglong),
                ("chainback_redo", ctypes.c_ulonglong),
                ("renormals", ctypes.c_longlong), ("stages", ctypes.c_longlong), ("walk_steps", ctypes.c_ulonglong)]


class MultiReport(ctypes.Structure):
    _fields_ = [("gpus", ctypes.c_int), ("handovers_verified", ctypes.c_int), ("ranges_redone", ctypes.c_int), ("worst_spread", ctypes.c_int),
                ("inner_verified", ctypes.c_int), ("inner_redone", ctypes.c_int), ("extra_stages", ctypes.c_longlong),
                ("residual_diffs", ctypes.c_longlong)]


class SegReport(ctypes.Structure):
    _fields_ = [("segments", ctypes.c_int), ("warm", ctypes.c_int), ("verified", ctypes.c_int), ("redone", ctypes.c_int),
                ("extra_stages", ctypes.c_longlong), ("worst_spread", ctypes.c_int)]


def library_path():
    # V224_LIB: an A/B build of the library (tools/build_variants.sh) instead of the in-tree one -- measurement runs only
    return os.environ.get("V224_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libviterbi224_b200.so")


_lib = None


def load_library():
    """Load the CUDA library.  Raises V224Error if it has not been built: nothing falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise V224Error(f"{path} is missing: build it with `python __graft_entry__.py build` (nvcc, sm_100a)")
    lib = ctypes.CDLL(path)
    vp, ci, cu, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_longlong
    sig = {
        "create_viterbi224": (vp, [ci]),
        "init_viterbi224": (ci, [vp, ci]),
        "update_viterbi224_blk": (ci, [vp, vp, ci]),
        "chainback_viterbi224": (ci, [vp, vp, cu, cu]),
        "decodebit_viterbi224": (ci, [vp, ci, ci]),
        "decodeword_viterbi224": (ctypes.c_ulonglong, [vp, ci, ci]),
        "max_metric_viterbi224": (ci, [vp]),
        "min_metric_viterbi224": (ci, [vp]),
        "delete_viterbi224": (None, [vp]),
        "v224x_device_count": (ci, []),
        "v224x_set_device": (ci, [ci]),
        "v224x_last_error": (ctypes.c_char_p, []),
        "v224x_version": (ctypes.c_char_p, []),
        "v224x_stream_decode": (ci, [vp, vp, ci, ci, vp]),
        "v224x_stream_decode_dev": (ci, [vp, vp, ci, ci, vp]),
        "v224x_stream_decode_seg": (ci, [vp, vp, ci, ci, vp, ci, ci, vp]),
        "v224x_stream_decode_seg_dev": (ci, [vp, vp, ci, ci, vp, ci, ci, vp]),
        "v224x_decode_frames": (ci, [vp, vp, ci, ci, vp, vp, vp, ci]),
        "v224x_update_dev": (ci, [vp, vp, ci]),
        "v224x_update_multi_dev": (ci, [vp, vp, ci, ci, vp]),
        "v224x_init_uniform": (ci, [vp, ci, ci]),
        "v224x_dev_alloc": (vp, [vp, ctypes.c_size_t]),
        "v224x_dev_free": (None, [vp, vp]),
        "v224x_h2d": (ci, [vp, vp, vp, ctypes.c_size_t]),
        "v224x_d2h": (ci, [vp, vp, vp, ctypes.c_size_t]),
        "v224x_host_alloc_pinned": (vp, [ctypes.c_size_t]),
        "v224x_host_free_pinned": (None, [vp]),
        "v224x_timer_start": (ci, [vp]),
        "v224x_timer_stop_ms": (ctypes.c_float, [vp]),
        "v224x_kernel_time_reset": (ci, [vp]),
        "v224x_kernel_time_enable": (ci, [vp, ci]),
        "v224x_kernel_time_ms": (ctypes.c_float, [vp, ctypes.POINTER(ctypes.c_ulonglong)]),
        "v224x_kernel_time_passes": (ctypes.c_ulonglong, [vp]),
        "v224x_get_stats": (ci, [vp, ctypes.POINTER(Stats)]),
        "v224x_get_metrics": (ci, [vp, vp]),
        "v224x_set_state": (ci, [vp, vp, cll, cll]),
        "v224x_get_row": (ci, [vp, ci, vp]),
        "v224x_set_option": (ci, [vp, ctypes.c_char_p, cll]),
        "v224x_pair_symbols": (cll, [vp, cll, ci, ci, ci, vp, vp, vp, ci, ctypes.POINTER(ci)]),
        "v224x_range_decode": (ci, [vp, vp, ci, ci, ci, vp, ci, ci, vp, vp, vp]),
        "v224x_range_decode_dev": (ci, [vp, vp, ci, ci, ci, vp, ci, ci, vp, vp, vp]),
        "v224x_metric_spread_dev": (ci, [vp, vp, vp, ctypes.POINTER(ci)]),
        "v224x_snapshot_bytes": (ctypes.c_size_t, []),
        "v224x_multi_create": (vp, [vp, ci, ci]),
        "v224x_multi_init": (ci, [vp, ci]),
        "v224x_multi_stream_decode": (ci, [vp, vp, cll, ci, vp, ci, ci, vp]),
        "v224x_multi_delete": (None, [vp]),
        "v224x_trim": (None, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def device_count():
    return load_library().v224x_device_count()


def pair_symbols(soft, start_phase=0, dontflip=False, delay=200, want_cmp=False):
    """v224x_pair_symbols: the pairs vdecode.c:101-140 hands to the decoder for the received symbols `soft` (host
    arithmetic, no GPU).  Returns (pairs uint8[n, 2], flips list) or (pairs, flips, cmp uint8[n, 2]) with want_cmp."""
    lib = load_library()
    a = np.ascontiguousarray(soft, dtype=np.uint8)
    cap = a.size // 2 + 1
    out = np.empty(2 * cap, dtype=np.uint8)
    cmp_ = np.empty(2 * cap, dtype=np.uint8) if want_cmp else None
    flips = np.zeros(4096, dtype=np.int64)
    nf = ctypes.c_int(0)
    n = lib.v224x_pair_symbols(a.ctypes.data_as(ctypes.c_void_p), a.size, int(start_phase), int(bool(dontflip)), int(delay),
                               out.ctypes.data_as(ctypes.c_void_p), None if cmp_ is None else cmp_.ctypes.data_as(ctypes.c_void_p),
                               flips.ctypes.data_as(ctypes.c_void_p), flips.size, ctypes.byref(nf))
    if n < 0:
        raise V224Error("v224x_pair_symbols failed")
    pairs = out[:2 * n].reshape(-1, 2)
    fl = [int(x) for x in flips[:min(nf.value, flips.size)]]
    return (pairs, fl, cmp_[:2 * n].reshape(-1, 2)) if want_cmp else (pairs, fl)


class MultiGpu:
    """v224x_multi_*: one stream decoded as time segments on several GPUs of this process (one host thread and one
    stream per GPU inside the library, hand-overs verified with peer copies of the metric snapshots)."""

    def __init__(self, ngpu, ring_rows, devices=None):
        self.lib = load_library()
        devs = None
        if devices is not None:
            devs = (ctypes.c_int * ngpu)(*devices)
        self.h = self.lib.v224x_multi_create(devs, int(ngpu), int(ring_rows))
        if not self.h:
            raise V224Error("v224x_multi_create failed: " + (self.lib.v224x_last_error() or b"").decode())

    def init(self, starting_state=0):
        if self.lib.v224x_multi_init(self.h, int(starting_state)) < 0:
            raise V224Error("v224x_multi_init failed: " + (self.lib.v224x_last_error() or b"").decode())

    def stream_decode(self, syms, delay, nseg=3, conv=-1, nbits=None, out=None):
        a, p = _u8(syms)
        n = a.size // 2 if nbits is None else int(nbits)
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        rep = MultiReport()
        rc = self.lib.v224x_multi_stream_decode(self.h, p, n, int(delay), out.ctypes.data_as(ctypes.c_void_p), int(nseg), int(conv), ctypes.byref(rep))
        if rc < 0:
            raise V224Error("v224x_multi_stream_decode failed: " + (self.lib.v224x_last_error() or b"").decode())
        return out, {k: getattr(rep, k) for k, _ in MultiReport._fields_}

    def delete(self):
        if getattr(self, "h", None):
            self.lib.v224x_multi_delete(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.delete()

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(ctypes.c_void_p)


class Viterbi224:
    """One decoder instance = one `void *` handle of the C ABI."""

    def __init__(self, length, device=None):
        self.lib = load_library()
        if device is not None and self.lib.v224x_set_device(int(device)) != 0:
            raise V224Error(self._err())
        self.len = int(length)
        self.h = self.lib.create_viterbi224(self.len)          # viterbi224.h:9
        if not self.h:
            raise V224Error("create_viterbi224 failed: " + self._err())

    def _err(self):
        return (self.lib.v224x_last_error() or b"").decode()

    def _check(self, rc, what):
        if rc < 0:
            raise V224Error(f"{what} failed: {self._err()}")
        return rc

    # ---- the nine reference entry points ----
    def init(self, starting_state=0):
        return self._check(self.lib.init_viterbi224(self.h, int(starting_state)), "init_viterbi224")

    def update_blk(self, syms, nbits=None):
        a, p = _u8(syms)
        n = a.size // 2 if nbits is None else int(nbits)
        assert a.size >= 2 * n
        return self._check(self.lib.update_viterbi224_blk(self.h, p, n), "update_viterbi224_blk")

    def chainback(self, nbits, endstate=0):
        out = np.zeros((int(nbits) + 7) // 8, dtype=np.uint8)
        self._check(self.lib.chainback_viterbi224(self.h, out.ctypes.data_as(ctypes.c_void_p), int(nbits), int(endstate) & 0xffffffff),
                    "chainback_viterbi224")
        return out

    def decodebit(self, delay, endstate=0):
        return self.lib.decodebit_viterbi224(self.h, int(delay), int(endstate))

    def decodeword(self, delay, endstate=0):
        return self.lib.decodeword_viterbi224(self.h, int(delay), int(endstate))

    def max_metric(self):
        return self.lib.max_metric_viterbi224(self.h)

    def min_metric(self):
        return self.lib.min_metric_viterbi224(self.h)

    def delete(self):
        if getattr(self, "h", None):
            self.lib.delete_viterbi224(self.h)
            self.h = None

    close = delete

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.delete()

    # ---- block-mode extensions ----
    def init_uniform(self, bias=5000, start_state=-1):
        return self._check(self.lib.v224x_init_uniform(self.h, int(bias), int(start_state)), "v224x_init_uniform")

    def stream_decode(self, syms, delay, nbits=None):
        """Block form of vdecode.c:145-152: returns (bits uint8[nbits], renormalisations)."""
        a, p = _u8(syms)
        n = a.size // 2 if nbits is None else int(nbits)
        out = np.empty(n, dtype=np.uint8)
        r = self._check(self.lib.v224x_stream_decode(self.h, p, n, int(delay), out.ctypes.data_as(ctypes.c_void_p)), "v224x_stream_decode")
        return out, r

    def stream_decode_seg(self, syms, delay, nseg, conv=-1, nbits=None):
        """v224x_stream_decode_seg: the stream in nseg segments advanced in lockstep, hand-overs verified on the device.
        Returns (bits uint8[nbits], report dict)."""
        a, p = _u8(syms)
        n = a.size // 2 if nbits is None else int(nbits)
        out = np.empty(n, dtype=np.uint8)
        rep = SegReport()
        self._check(self.lib.v224x_stream_decode_seg(self.h, p, n, int(delay), out.ctypes.data_as(ctypes.c_void_p), int(nseg), int(conv),
                                                     ctypes.byref(rep)), "v224x_stream_decode_seg")
        return out, {k: getattr(rep, k) for k, _ in SegReport._fields_}

    def stream_decode_seg_dev(self, dev_syms, nbits, delay, dev_bits, nseg, conv=-1):
        rep = SegReport()
        self._check(self.lib.v224x_stream_decode_seg_dev(self.h, dev_syms, int(nbits), int(delay), dev_bits, int(nseg), int(conv),
                                                         ctypes.byref(rep)), "v224x_stream_decode_seg_dev")
        return {k: getattr(rep, k) for k, _ in SegReport._fields_}

    def range_decode(self, syms, lead, nout, delay, nseg=3, conv=-1, snap_early=None, snap_late=None):
        """v224x_range_decode: one time segment of a longer stream (host buffers).  snap_early / snap_late: device
        pointers (16 MiB each) or None.  Returns (bits uint8[nout], report dict)."""
        a, p = _u8(syms)
        assert a.size >= 2 * (lead + nout)
        out = np.empty(nout, dtype=np.uint8)
        rep = SegReport()
        self._check(self.lib.v224x_range_decode(self.h, p, int(lead), int(nout), int(delay), out.ctypes.data_as(ctypes.c_void_p), int(nseg),
                                                int(conv), snap_early, snap_late, ctypes.byref(rep)), "v224x_range_decode")
        return out, {k: getattr(rep, k) for k, _ in SegReport._fields_}

    def range_decode_dev(self, dev_syms, lead, nout, delay, dev_bits, nseg=3, conv=-1, snap_early=None, snap_late=None):
        rep = SegReport()
        self._check(self.lib.v224x_range_decode_dev(self.h, dev_syms, int(lead), int(nout), int(delay), dev_bits, int(nseg), int(conv),
                                                    snap_early, snap_late, ctypes.byref(rep)), "v224x_range_decode_dev")
        return {k: getattr(rep, k) for k, _ in SegReport._fields_}

    def metric_spread_dev(self, dev_a, dev_b):
        """0 <=> the two metric snapshots differ by a constant (the two decoders are in step)."""
        out = ctypes.c_int(0)
        self._check(self.lib.v224x_metric_spread_dev(self.h, dev_a, dev_b, ctypes.byref(out)), "v224x_metric_spread_dev")
        return int(out.value)

    def decode_frames(self, syms, nframes, framebits, start_states=None, end_states=None, nlock=3):
        """v224x_decode_frames: nframes independent frames (init / update / chainback each), nlock side by side.
        Returns uint8[nframes, ceil(framebits/8)]."""
        a, p = _u8(syms)
        assert a.size >= 2 * nframes * framebits
        out = np.empty((nframes, (framebits + 7) // 8), dtype=np.uint8)
        ss = None if start_states is None else np.ascontiguousarray(start_states, dtype=np.uint32)
        es = None if end_states is None else np.ascontiguousarray(end_states, dtype=np.uint32)
        self._check(self.lib.v224x_decode_frames(self.h, p, int(nframes), int(framebits),
                                                 None if ss is None else ss.ctypes.data_as(ctypes.c_void_p),
                                                 None if es is None else es.ctypes.data_as(ctypes.c_void_p),
                                                 out.ctypes.data_as(ctypes.c_void_p), int(nlock)), "v224x_decode_frames")
        return out

    def dev_alloc(self, nbytes):
        p = self.lib.v224x_dev_alloc(self.h, int(nbytes))
        if not p:
            raise V224Error("v224x_dev_alloc failed: " + self._err())
        return p

    def dev_free(self, ptr):
        self.lib.v224x_dev_free(self.h, ptr)

    def h2d(self, dev_ptr, host_array):
        a = np.ascontiguousarray(host_array)
        self._check(self.lib.v224x_h2d(self.h, dev_ptr, a.ctypes.data_as(ctypes.c_void_p), a.nbytes), "v224x_h2d")

    def d2h(self, host_array, dev_ptr):
        assert host_array.flags["C_CONTIGUOUS"]
        self._check(self.lib.v224x_d2h(self.h, host_array.ctypes.data_as(ctypes.c_void_p), dev_ptr, host_array.nbytes), "v224x_d2h")

    def update_dev(self, dev_syms, nbits):
        return self._check(self.lib.v224x_update_dev(self.h, dev_syms, int(nbits)), "v224x_update_dev")

    @staticmethod
    def update_multi_dev(decoders, dev_syms, nbits):
        """Advance several decoders of one GPU in lockstep (v224x_update_multi_dev).  Returns per-decoder renorm counts."""
        lib = load_library()
        n = len(decoders)
        hs = (ctypes.c_void_p * n)(*[d.h for d in decoders])
        ps = (ctypes.c_void_p * n)(*dev_syms)
        ren = (ctypes.c_int * n)()
        rc = lib.v224x_update_multi_dev(hs, ps, n, int(nbits), ren)
        if rc < 0:
            raise V224Error("v224x_update_multi_dev failed: " + (lib.v224x_last_error() or b"").decode())
        return list(ren)

    def stream_decode_dev(self, dev_syms, nbits, delay, dev_bits):
        return self._check(self.lib.v224x_stream_decode_dev(self.h, dev_syms, int(nbits), int(delay), dev_bits), "v224x_stream_decode_dev")

    def timer_start(self):
        self._check(self.lib.v224x_timer_start(self.h), "v224x_timer_start")

    def timer_stop_ms(self):
        return float(self.lib.v224x_timer_stop_ms(self.h))

    def kernel_time_enable(self, on=True):
        self.lib.v224x_kernel_time_enable(self.h, 1 if on else 0)
        self.lib.v224x_kernel_time_reset(self.h)

    def kernel_time_ms(self):
        n = ctypes.c_ulonglong(0)
        ms = float(self.lib.v224x_kernel_time_ms(self.h, ctypes.byref(n)))
        return ms, int(n.value), int(self.lib.v224x_kernel_time_passes(self.h))

    def stats(self):
        s = Stats()
        self._check(self.lib.v224x_get_stats(self.h, ctypes.byref(s)), "v224x_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    # ---- test hooks ----
    def get_metrics(self):
        out = np.empty(NSTATES, dtype=np.int16)
        self._check(self.lib.v224x_get_metrics(self.h, out.ctypes.data_as(ctypes.c_void_p)), "v224x_get_metrics")
        return out

    def set_state(self, metrics, renormals=0, stages=0):
        m = np.ascontiguousarray(metrics, dtype=np.int16)
        assert m.size == NSTATES
        self._check(self.lib.v224x_set_state(self.h, m.ctypes.data_as(ctypes.c_void_p), int(renormals), int(stages)), "v224x_set_state")

    def get_row(self, row):
        out = np.empty(ROWWORDS, dtype=np.uint32)
        self._check(self.lib.v224x_get_row(self.h, int(row), out.ctypes.data_as(ctypes.c_void_p)), "v224x_get_row")
        return out

    def set_option(self, key, value):
        self._check(self.lib.v224x_set_option(self.h, key.encode(), int(value)), "v224x_set_option")
