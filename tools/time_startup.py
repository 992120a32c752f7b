"""Where the start-up time of a frame program goes (fresh process): CUDA context, first decoder, further decoders, first batch.
usage: time_startup.py [framebits]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.perf_counter()
import numpy as np
import isee3_decoder_b200 as v224
lib = v224.load_library()
t1 = time.perf_counter()
fb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = lib.v224x_device_count()
t2 = time.perf_counter()
d = v224.Viterbi224(fb)
t3 = time.perf_counter()
d2 = v224.Viterbi224(fb)
t4 = time.perf_counter()
d3 = v224.Viterbi224(fb)
t5 = time.perf_counter()
nframes = 64
syms = np.full(2 * fb * nframes, 128, np.uint8)
d.decode_frames(syms, 8, fb, nlock=4)
t6 = time.perf_counter()
d.decode_frames(syms, nframes, fb, nlock=4)
t7 = time.perf_counter()
d.decode_frames(syms, nframes, fb, nlock=4)
t8 = time.perf_counter()
print(f"import+dlopen {t1 - t0:.3f} s | device count ({n}) {t2 - t1:.3f} s | first create({fb}) incl. CUDA context {t3 - t2:.3f} s | second create {t4 - t3:.3f} s | "
      f"third create {t5 - t4:.3f} s | first decode_frames(8 frames, creates 7 more decoders) {t6 - t5:.3f} s | decode_frames({nframes}) {t7 - t6:.3f} s | again {t8 - t7:.3f} s "
      f"= {nframes / (t8 - t7):.0f} frames/s")
