"""Cost of create_viterbi224 / delete_viterbi224 (decode.c:216-229 creates and deletes a 1024-row decoder per frame)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isee3_decoder_b200 as v224
d = v224.Viterbi224(1024); d.delete()
for n in (1024, 201, 8392):
    t0 = time.perf_counter()
    for _ in range(20):
        d = v224.Viterbi224(n); d.delete()
    dt = (time.perf_counter() - t0) / 20
    print(f"create({n}) + delete: {1e3 * dt:.2f} ms")
