"""TEST INFRASTRUCTURE ONLY -- ctypes front ends for the CPU oracle (oracle/v224_oracle.c, our
restatement) and, where oracle/_ref exists, the unmodified reference compiled from
/root/reference (oracle/Makefile).  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NSTATES = 1 << 23
ROWWORDS = 1 << 18
ORACLE_SO = os.path.join(HERE, "_build", "libv224_oracle.so")
REF_SSE2_SO = os.path.join(HERE, "_ref", "libv224_sse2.so")
REF_PORT_SO = os.path.join(HERE, "_ref", "libv224_port.so")
REF_UTIL_SO = os.path.join(HERE, "_ref", "libv224_refutil.so")
REF_VDECODE = os.path.join(HERE, "_ref", "vdecode_sse")
REF_VTEST = os.path.join(HERE, "_ref", "vtest224sse")


def build(ref=True):
    """make the oracle (always) and the reference objects (only where /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref():
    return os.path.exists(REF_SSE2_SO)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _Base:
    """Common surface: the nine ABI calls + inspection, over different C symbol names."""
    kind = "base"

    def update_blk(self, syms, nbits=None):
        a = np.ascontiguousarray(syms, dtype=np.uint8)
        n = a.size // 2 if nbits is None else int(nbits)
        return self._update(a, n)

    def stream_decode(self, syms, delay, nbits=None):
        """The vdecode.c:145-152 loop, literally: update(1) + decodebit(delay, 0) per bit."""
        a = np.ascontiguousarray(syms, dtype=np.uint8)
        n = a.size // 2 if nbits is None else int(nbits)
        out = np.empty(n, dtype=np.uint8)
        ren = 0
        for i in range(n):
            ren += self._update(a[2 * i: 2 * i + 2], 1)
            out[i] = self.decodebit(delay, 0) & 0xFF
        return out, ren

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.delete()


class Oracle(_Base):
    """oracle/v224_oracle.c"""
    kind = "port"
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build(ref=False)
            L = ctypes.CDLL(ORACLE_SO)
            L.oracle_create.restype = ctypes.c_void_p
            L.oracle_create.argtypes = [ctypes.c_int]
            for name, res, args in [
                ("oracle_init", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
                ("oracle_delete", None, [ctypes.c_void_p]),
                ("oracle_update_blk", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
                ("oracle_chainback", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]),
                ("oracle_decodebit", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
                ("oracle_decodeword", ctypes.c_ulonglong, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
                ("oracle_max_metric", ctypes.c_int, [ctypes.c_void_p]),
                ("oracle_min_metric", ctypes.c_int, [ctypes.c_void_p]),
                ("oracle_metrics", ctypes.POINTER(ctypes.c_int16), [ctypes.c_void_p]),
                ("oracle_metrics_mut", ctypes.POINTER(ctypes.c_int16), [ctypes.c_void_p]),
                ("oracle_row", ctypes.POINTER(ctypes.c_uint32), [ctypes.c_void_p, ctypes.c_longlong]),
                ("oracle_renormals", ctypes.c_longlong, [ctypes.c_void_p]),
                ("oracle_set_renormals", None, [ctypes.c_void_p, ctypes.c_longlong]),
                ("oracle_dp", ctypes.c_longlong, [ctypes.c_void_p]),
                ("oracle_set_dp", None, [ctypes.c_void_p, ctypes.c_longlong]),
                ("oracle_encode", ctypes.c_ulonglong, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_ulonglong]),
                ("oracle_setup_channel", None, [ctypes.c_double, ctypes.c_double]),
                ("oracle_simulate", ctypes.c_ubyte, [ctypes.c_int]),
                ("oracle_simulate_draw", ctypes.c_ubyte, [ctypes.c_int, ctypes.c_int]),
                ("oracle_srandom", None, [ctypes.c_uint]),
            ]:
                f = getattr(L, name)
                f.restype = res
                f.argtypes = args
            cls._lib = L
        return cls._lib

    def __init__(self, length):
        self.L = self.lib()
        self.len = int(length)
        self.h = self.L.oracle_create(self.len)
        if not self.h:
            raise MemoryError("oracle_create failed")

    def init(self, starting_state=0):
        return self.L.oracle_init(self.h, int(starting_state))

    def _update(self, a, n):
        return self.L.oracle_update_blk(self.h, _p(a), n)

    def chainback(self, nbits, endstate=0):
        out = np.zeros((int(nbits) + 7) // 8, dtype=np.uint8)
        self.L.oracle_chainback(self.h, _p(out), int(nbits), int(endstate) & 0xFFFFFFFF)
        return out

    def decodebit(self, delay, endstate=0):
        return self.L.oracle_decodebit(self.h, int(delay), int(endstate))

    def decodeword(self, delay, endstate=0):
        return self.L.oracle_decodeword(self.h, int(delay), int(endstate))

    def max_metric(self):
        return self.L.oracle_max_metric(self.h)

    def min_metric(self):
        return self.L.oracle_min_metric(self.h)

    def get_metrics(self):
        return np.ctypeslib.as_array(self.L.oracle_metrics(self.h), (NSTATES,)).copy()

    def set_state(self, metrics, renormals=0, stages=0):
        dst = np.ctypeslib.as_array(self.L.oracle_metrics_mut(self.h), (NSTATES,))
        dst[:] = np.asarray(metrics, dtype=np.int16)
        self.L.oracle_set_renormals(self.h, int(renormals))
        self.L.oracle_set_dp(self.h, int(stages))

    def get_row(self, row):
        return np.ctypeslib.as_array(self.L.oracle_row(self.h, int(row)), (ROWWORDS,)).copy()

    def renormals(self):
        return self.L.oracle_renormals(self.h)

    def delete(self):
        if self.h:
            self.L.oracle_delete(self.h)
            self.h = None


class RefSSE2(_Base):
    """The unmodified viterbi224_sse2.c (+ accessors of oracle/ref_shim_sse2.c)."""
    kind = "reference"
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = ctypes.CDLL(REF_SSE2_SO)
            vp, ci = ctypes.c_void_p, ctypes.c_int
            for name, res, args in [
                ("create_viterbi224", vp, [ci]), ("init_viterbi224", ci, [vp, ci]),
                ("update_viterbi224_blk", ci, [vp, vp, ci]),
                ("chainback_viterbi224", ci, [vp, vp, ctypes.c_uint, ctypes.c_uint]),
                ("decodebit_viterbi224", ci, [vp, ci, ci]), ("decodeword_viterbi224", ctypes.c_ulonglong, [vp, ci, ci]),
                ("max_metric_viterbi224", ci, [vp]), ("min_metric_viterbi224", ci, [vp]), ("delete_viterbi224", None, [vp]),
                ("refshim_metrics", ctypes.POINTER(ctypes.c_int16), [vp]), ("refshim_metrics_mut", ctypes.POINTER(ctypes.c_int16), [vp]),
                ("refshim_row", ctypes.POINTER(ctypes.c_uint32), [vp, ci]), ("refshim_renormals", ctypes.c_longlong, [vp]),
                ("refshim_set_renormals", None, [vp, ctypes.c_longlong]), ("refshim_dp", ci, [vp]), ("refshim_set_dp", None, [vp, ci]),
                ("refshim_zero_ring", None, [vp]),
            ]:
                f = getattr(L, name)
                f.restype = res
                f.argtypes = args
            cls._lib = L
        return cls._lib

    def __init__(self, length):
        self.L = self.lib()
        self.len = int(length)
        self.h = self.L.create_viterbi224(self.len)
        if not self.h:
            raise MemoryError("create_viterbi224 (reference) failed")
        self.L.refshim_zero_ring(self.h)

    def init(self, starting_state=0):
        return self.L.init_viterbi224(self.h, int(starting_state))

    def _update(self, a, n):
        return self.L.update_viterbi224_blk(self.h, _p(a), n)

    def chainback(self, nbits, endstate=0):
        out = np.zeros((int(nbits) + 7) // 8, dtype=np.uint8)
        self.L.chainback_viterbi224(self.h, _p(out), int(nbits), int(endstate) & 0xFFFFFFFF)
        return out

    def decodebit(self, delay, endstate=0):
        return self.L.decodebit_viterbi224(self.h, int(delay), int(endstate))

    def decodeword(self, delay, endstate=0):
        return self.L.decodeword_viterbi224(self.h, int(delay), int(endstate))

    def max_metric(self):
        return self.L.max_metric_viterbi224(self.h)

    def min_metric(self):
        return self.L.min_metric_viterbi224(self.h)

    def get_metrics(self):
        return np.ctypeslib.as_array(self.L.refshim_metrics(self.h), (NSTATES,)).copy()

    def set_state(self, metrics, renormals=0, stages=0):
        dst = np.ctypeslib.as_array(self.L.refshim_metrics_mut(self.h), (NSTATES,))
        dst[:] = np.asarray(metrics, dtype=np.int16)
        self.L.refshim_set_renormals(self.h, int(renormals))
        self.L.refshim_set_dp(self.h, int(stages) % self.len)

    def get_row(self, row):
        return np.ctypeslib.as_array(self.L.refshim_row(self.h, int(row)), (ROWWORDS,)).copy()

    def renormals(self):
        return self.L.refshim_renormals(self.h)

    def delete(self):
        if self.h:
            self.L.delete_viterbi224(self.h)
            self.h = None


class RefPort(_Base):
    """The unmodified viterbi224_port.c (secondary oracle: different tie-break/bias, no renorm)."""
    kind = "reference-port"
    _lib = None

    def __init__(self, length):
        if RefPort._lib is None:
            L = ctypes.CDLL(REF_PORT_SO)
            vp, ci = ctypes.c_void_p, ctypes.c_int
            L.create_viterbi224.restype = vp
            L.create_viterbi224.argtypes = [ci]
            L.init_viterbi224.argtypes = [vp, ci]
            L.update_viterbi224_blk.argtypes = [vp, vp, ci]
            L.chainback_viterbi224.argtypes = [vp, vp, ctypes.c_uint, ctypes.c_uint]
            L.decodebit_viterbi224.argtypes = [vp, ci, ci]
            L.delete_viterbi224.argtypes = [vp]
            L.delete_viterbi224.restype = None
            RefPort._lib = L
        self.L = RefPort._lib
        self.len = int(length)
        self.h = self.L.create_viterbi224(self.len)

    def init(self, starting_state=0):
        return self.L.init_viterbi224(self.h, int(starting_state))

    def _update(self, a, n):
        return self.L.update_viterbi224_blk(self.h, _p(a), n)

    def chainback(self, nbits, endstate=0):
        out = np.zeros((int(nbits) + 7) // 8, dtype=np.uint8)
        self.L.chainback_viterbi224(self.h, _p(out), int(nbits), int(endstate) & 0xFFFFFFFF)
        return out

    def decodebit(self, delay, endstate=0):
        return self.L.decodebit_viterbi224(self.h, int(delay), int(endstate))

    def delete(self):
        if self.h:
            self.L.delete_viterbi224(self.h)
            self.h = None


def best_cpu_decoder():
    """The strongest CPU checker available: the real reference if it was compiled here, else our port."""
    return RefSSE2 if have_ref() else Oracle


def row_crc(row):
    import zlib
    return zlib.crc32(np.ascontiguousarray(row).tobytes()) & 0xFFFFFFFF
