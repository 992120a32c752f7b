"""Build libviterbi224_b200.so (sm_100a) in-tree with nvcc.  No torch involved: the library is
a plain C-ABI shared object so that the reference's C callers can link against it."""
import os
import subprocess
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libviterbi224_b200.so")
# (source, extra nvcc flags).  The fused ACS pass is compiled with ptxas -O1: at the default level ptxas reorders the
# decision-bit gather behind a whole stage of butterflies and spills; in source order the tile body needs no spill.
# The fused pass is built twice: 64-column tiles (lockstep decoders) and 32-column tiles (a decoder running alone).
SOURCES = [("v224_acs_persist.cu", ["-Xptxas", "-O1"], "v224_acs_persist.o"),
           ("v224_acs_persist.cu", ["-Xptxas", "-O1", "-DV224_TILE_COLS_LOG2=5", "-DV224_NS=v224t32", "-DV224_BRIDGE=v224_t32"], "v224_acs_persist_t32.o"),

           ("v224_kernels.cu", [], "v224_kernels.o"), ("v224_runtime.cu", [], "v224_runtime.o")]
# A/B only (build.py --out ... -DV224_WITH_Q1): 32-column tiles with ONE packed register per row and thread (256 threads on half a
# tile's work).  Correct (emulated on the CPU tier, all golden variants on the GPU) but slower than the two-register build for a lone
# decoder (13.9 against 12.9 us per pass) and in lockstep (10.5 against 9.1): profiles/r02_probe_one_register_build.txt.
Q1_SOURCE = ("v224_acs_persist.cu", ["-Xptxas", "-O1", "-DV224_TILE_COLS_LOG2=5", "-DV224_NQ=1", "-DV224_NS=v224t32q1", "-DV224_BRIDGE=v224_t32q1"],
             "v224_acs_persist_t32q1.o")
HOST_SOURCES = ["v224_pairing.cpp"]           # plain C++ (g++ -O3): host-side entries of the library, no CUDA
HEADERS = ["v224_common.cuh", "v224_fused_core.cuh", "v224_kernels.h", "v224_pairing.cpp", os.path.join("..", "host", "pairing.h"),
           os.path.join("..", "..", "include", "viterbi224.h"), os.path.join("..", "..", "include", "viterbi224_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--cudart", "static", "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in [x[0] for x in SOURCES] + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


HOST_PROGRAMS = {"vdecode_block": os.path.join(HERE, "host", "vdecode_block.cpp"),
                 "decode_block": os.path.join(HERE, "host", "decode_block.cpp")}
HOST_HEADERS = [os.path.join(HERE, "host", "hostfmt.h"), os.path.join(HERE, "host", "fano_seq.h"), os.path.join(HERE, "host", "pairing.h")]
BIN = os.path.join(HERE, "bin")


def build_host_programs(force=False):
    """Host-side programs above the C ABI (plain C++, no CUDA): linked against the in-tree library, rpath relative."""
    os.makedirs(BIN, exist_ok=True)
    outs = []
    for name, src in HOST_PROGRAMS.items():
        exe = os.path.join(BIN, name)
        if force or not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(LIB), *map(os.path.getmtime, HOST_HEADERS)):
            cmd = ["g++", "-O3", "-Wall", "-std=c++17", "-o", exe, src, "-L" + HERE, "-lviterbi224_b200", "-Wl,-rpath,$ORIGIN/.."]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("host program build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        outs.append(exe)
    return outs


def build_library(force=False, verbose=False, out=None, extra_flags=()):
    """Compile the CUDA sources for sm_100a and link the shared library.  Returns its path.
    out / extra_flags: A/B builds of kernel-shape variants into another file (tools/build_variants.sh)."""
    variant = out is not None
    if not variant and not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    objs = []
    log = []
    for s, flags, oname in SOURCES + ([Q1_SOURCE] if "-DV224_WITH_Q1" in extra_flags else []):
        o = (out + "." if variant else os.path.join(CSRC, "")) + oname
        cmd = [nvcc, *NVCC_FLAGS, *flags, *extra_flags, "-c", "-o", o, os.path.join(CSRC, s)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        objs.append(o)
    for s in HOST_SOURCES:
        o = (out + "." if variant else os.path.join(CSRC, "")) + s.replace(".cpp", ".o")
        cmd = ["g++", "-O3", "-std=c++17", "-Wall", "-fPIC", "-c", "-o", o, os.path.join(CSRC, s)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        objs.append(o)
    lib = out if variant else LIB
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(lib + ".ptxas.log" if variant else os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(l for l in "\n".join(log).split("\n") if "Compile time" not in l))
    if variant:
        for o in objs:
            os.remove(o)
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "--out":
        print(build_library(out=os.path.abspath(sys.argv[2]), extra_flags=sys.argv[3:]))
    else:
        print(build_library(force=True, verbose=True))
        print(build_host_programs(force=True))
