"""Frames per second of the frame decoder programs on one stream of clean-locking minor frames:
   decode_block -V (speculative batches, frames side by side)   vs   stock decode.c -V linked against the same library
   (one init / update / chainback per frame through the ABI, oracle/_ref/decode_b200).  Output must be identical."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isee3_decoder_b200 as v224

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
_, soft = v224.streams.telemetry_stream((nframes + 1) * 1024, 3.0, seed=4242, junk_symbols=500)
env = dict(os.environ, LANG="C", V224_HOST_STATS="1")
blk = os.path.join(ROOT, "isee3-decoder_b200", "bin", "decode_block")
stock = os.path.join(ROOT, "oracle", "_ref", "decode_b200")
res = {}
for name, cmd, frames in (("decode_block -V", [blk, "-V"], nframes), ("decode_block -V -B 256", [blk, "-V", "-B", "256"], nframes),
                          ("stock decode -V on libviterbi224_b200", [stock, "-V"], min(nframes, 256))):
    if not os.path.exists(cmd[0]):
        print(f"{name}: not built"); continue
    data = soft[: 500 + (frames + 1) * 2048].tobytes()
    t0 = time.perf_counter()
    r = subprocess.run(cmd, input=data, capture_output=True, env=env)
    dt = time.perf_counter() - t0
    n = r.stdout.count(b"Frame ")
    res[name] = r.stdout.split(b"\n", 1)[1][: 400 * 200]
    print(f"{name:42s} {n:5d} frames ({r.stdout.count(b'(bad)')} bad) in {dt:6.2f} s = {n / dt:7.1f} frames/s (process start and create included)  {r.stderr.decode().strip()[-230:]}")
vals = list(res.values())
print("outputs identical over the common prefix:", all(v[: min(map(len, vals))] == vals[0][: min(map(len, vals))] for v in vals))

# Fano first at an Eb/N0 where the sequential decoder gives up on a few percent of the frames (they go to the GPU in batches)
nf2 = min(nframes, 512)
_, soft2 = v224.streams.telemetry_stream((nf2 + 1) * 1024, 1.9, seed=4343, junk_symbols=500)
res2 = {}
for name, cmd in (("decode_block (Fano first, 1.9 dB)", [blk]), ("stock decode (Fano first) on libviterbi224_b200", [stock])):
    if not os.path.exists(cmd[0]):
        print(f"{name}: not built"); continue
    t0 = time.perf_counter()
    r = subprocess.run(cmd, input=soft2.tobytes(), capture_output=True, env=env)
    dt = time.perf_counter() - t0
    n = r.stdout.count(b"Frame ")
    res2[name] = r.stdout.split(b"\n", 2)[2]
    print(f"{name:48s} {n:5d} frames ({r.stdout.count(b'with Viterbi')} by Viterbi, {r.stdout.count(b'(bad)')} bad) in {dt:6.2f} s = {n / dt:7.1f} frames/s  {r.stderr.decode().strip()[-230:]}")
print("outputs identical:", len(set(res2.values())) == 1)
