#!/usr/bin/env python
"""Regenerates tests/golden/host/hybridtest_seed20141_1dB.txt: the printout of the UNMODIFIED hybridtest.c on the reference's
SSE2 decoder (oracle/_ref/hybridtest_sse, built by oracle/Makefile) for 30 frames at Eb/N0 = 1.0 dB, with its seed pinned
(hybridtest.c:114 seeds with time(); oracle/fixed_time_shim.c, preloaded, makes time() return V224_FIXED_TIME).
15 of the 30 frames go to the Viterbi decoder (hybridtest.c:186-193: create / init / update / chainback / delete per frame),
4 of them decode with errors -- the GPU tier runs the same unchanged program linked against libviterbi224_b200.so
(oracle/_ref/hybridtest_b200) under the same shim and compares every line.  About one CPU-minute.

    python tools/make_golden_hybridtest.py          (needs /root/reference, i.e. oracle/_ref built here)
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
ARGS = ["-n", "30", "-e", "1.0", "-v"]
SEED = "20141"


def run(exe):
    env = dict(os.environ, LD_PRELOAD=os.path.join(REF, "libfixed_time.so"), V224_FIXED_TIME=SEED)
    return subprocess.run([os.path.join(REF, exe)] + ARGS, capture_output=True, text=True, env=env, check=True).stdout


if __name__ == "__main__":
    out = run("hybridtest_sse")
    dst = os.path.join(ROOT, "tests", "golden", "host", "hybridtest_seed20141_1dB.txt")
    open(dst, "w").write(out)
    print(dst)
    print("\n".join(out.splitlines()[-2:]))
