"""How often the speculative chainback has to re-walk a segment (counter chainback_redo): BASELINE config 1 frames (10,000
bits, vtest-style AWGN) at several Eb/N0, default segment length / warm-up and two shorter settings.  The result always
equals the serial walk (a wrong guess is detected against the neighbour's arrival state and the segment is walked again);
the rate only costs latency.  usage: report_chainback_redo.py [frames per point]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import isee3_decoder_b200 as v224
nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fb = 10000
with v224.Viterbi224(fb) as d:
    for seg, warm in ((128, 256), (64, 192), (32, 224), (32, 128)):
        d.set_option("chain_seg", seg)
        d.set_option("chain_warm", warm)
        for ebn0 in (3.0, 2.0, 1.5, 1.0):
            redo0 = d.stats()["chainback_redo"]
            errs = 0
            t_cb = 0.0
            for f in range(nfr):
                data, syms = v224.streams.vtest_frame(fb, ebn0, seed=1000 * f + int(10 * ebn0))
                d.init(0)
                d.update_blk(syms, fb)
                t0 = time.perf_counter()
                out = d.chainback(fb, 0)
                t_cb += time.perf_counter() - t0
                errs += int(np.unpackbits(out ^ data).sum())
            redo = d.stats()["chainback_redo"] - redo0
            nseg = nfr * ((fb + seg - 1) // seg)
            print(f"chain_seg {seg:4d} chain_warm {warm:4d}  Eb/N0 {ebn0:3.1f} dB: {redo:5d} of {nseg:6d} segments re-walked ({100.0 * redo / nseg:6.3f} %), "
                  f"chainback {1e3 * t_cb / nfr:6.2f} ms per frame, bit errors {errs} of {nfr * fb}", flush=True)
