"""Build libviterbi224_b200.so (sm_100a) in-tree with nvcc.  No torch involved: the library is
a plain C-ABI shared object so that the reference's C callers can link against it."""
import os
import subprocess
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libviterbi224_b200.so")
SOURCES = ["v224_kernels.cu", "v224_runtime.cu"]
HEADERS = ["v224_common.cuh", "v224_fused_core.cuh", "v224_kernels.h",
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
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile the CUDA sources for sm_100a and link the shared library.  Returns its path."""
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    objs = []
    log = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", o, os.path.join(CSRC, s)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        objs.append(o)
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
