"""isee3-decoder_b200 -- B200-native K=24 r=1/2 Viterbi decoder (viterbi224) behind the
reference's libfec-style C ABI.

The product is the C-ABI shared library ``libviterbi224_b200.so`` (sources in ``csrc/``,
headers in ``/include``).  This package is the thin Python host side used by the tests, the
benchmark and Python callers: a ctypes binding that mirrors the nine reference entry points
(viterbi224.h:8-16) name for name, the block-mode extensions, the synthetic stream generators
(the way vtest224.c / sim.c / symdemod.c produce decoder input) and a block-mode mirror of the
reference's streaming driver vdecode.c.

There is no CPU decoding path: loading fails loudly if the CUDA library is missing, and
``Viterbi224(...)`` raises if no CUDA device is usable.
"""
from .binding import (Viterbi224, MultiGpu, pair_symbols, V224Error, load_library, library_path, device_count, NSTATES, ROWWORDS,  # noqa: F401
                      ABI_SYMBOLS, EXT_SYMBOLS)
from . import streams  # noqa: F401
from . import vdecode  # noqa: F401
from . import segments  # noqa: F401
