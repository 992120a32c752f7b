"""CPU tier: the C-ABI library loads, exports every symbol declared in include/*.h, and follows the
reference's error conventions for NULL handles (viterbi224_sse2.c:41-42,87-88,120-121,169-170).
No compute is attempted without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

import isee3_decoder_b200 as v224

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(\w+)\s*\([^;{]*\)\s*;", src)


def test_headers_declare_the_reference_abi():
    names = declared_functions("viterbi224.h")
    assert sorted(names) == sorted(v224.ABI_SYMBOLS)     # exactly the nine of viterbi224.h:8-16
    ext = declared_functions("viterbi224_b200.h")
    assert sorted(ext) == sorted(v224.EXT_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", v224.library_path()], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in v224.ABI_SYMBOLS + v224.EXT_SYMBOLS:
        assert name in exported, f"{name} not exported"
    # C linkage: no mangled product symbols leak as the ABI
    lib = v224.load_library()
    for name in v224.ABI_SYMBOLS + v224.EXT_SYMBOLS:
        assert getattr(lib, name) is not None


def test_library_contains_sm100a_code(built):
    out = subprocess.run(["cuobjdump", "--list-elf", v224.library_path()], capture_output=True, text=True)
    assert "sm_100a" in out.stdout


def test_fused_pass_is_built_from_the_instructions_the_design_names(built):
    """DESIGN.md section 3.1 rests on these SASS instructions being what the fused pass compiles to (both tile shapes): the tile's
    2-D tensor copy and the pass table's bulk copy (TMA), mbarrier waits, the packed add-min of the butterfly, three-input adds
    with a negated operand for the pairs that balance the integer pipes, 256-bit metric stores, streaming decision stores, and
    (64-column build) the L2 discard of consumed input lines."""
    out = subprocess.run(["cuobjdump", "-sass", v224.library_path()], capture_output=True, text=True).stdout
    body = {}
    name = None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            body[name] = []
        elif name is not None:
            body[name].append(line)
    fused = {n: "\n".join(b) for n, b in body.items() if "k_acs_persist" in n}
    assert len(fused) == 2, sorted(body)                                   # v224:: (64 columns) and v224t32:: (32 columns)
    for n, text in fused.items():
        for needle in ("UTMALDG.2D", "UBLKCP", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "VIADDMNMX.U16x2", "STG.E.ENL2.256", "STG.E.EF"):
            assert needle in text, (n, needle)
        assert text.count("VIADDMNMX.U16x2") == 512, n                   # 32 per stage x 8 stages x {careful, plain} tile bodies
        assert len(re.findall(r"IADD3 R\d+, PT, PT, [^;]*-R\d+", text)) >= 64, n       # c - a + K, 4 per stage x 8 x 2
        assert "LDL" not in text and "STL" not in text, n                 # no spill in either build (csrc/ptxas.log)
    assert "CCTL.E.RML2" in next(t for n, t in fused.items() if "t32" not in n)


def test_null_handle_conventions(built):
    lib = v224.load_library()
    buf = ctypes.create_string_buffer(16)
    assert lib.init_viterbi224(None, 0) == -1
    assert lib.update_viterbi224_blk(None, buf, 1) == -1
    assert lib.chainback_viterbi224(None, buf, 8, 0) == -1
    assert lib.decodebit_viterbi224(None, 10, 0) == -1
    assert lib.max_metric_viterbi224(None) == -1
    assert lib.min_metric_viterbi224(None) == -1
    lib.delete_viterbi224(None)            # NULL-safe, viterbi224_sse2.c:251
    assert lib.v224x_stream_decode(None, buf, 1, 1, buf) == -1


def test_create_rejects_bad_length(built):
    lib = v224.load_library()
    assert not lib.create_viterbi224(0)
    assert not lib.create_viterbi224(-5)


def test_no_cpu_fallback_without_gpu(built):
    """Without a CUDA device the product refuses to create a decoder (it must not decode on the CPU)."""
    if v224.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(v224.V224Error):
        v224.Viterbi224(8)
    assert b"no CUDA device" in v224.load_library().v224x_last_error()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "isee3-decoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "v224_oracle" not in text and "oracle/" not in text, f
    needed = subprocess.run(["ldd", v224.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in needed


def test_reference_callers_link_unchanged():
    """tools/build_dropin.sh: vtest224.c / vdecode.c / hybridtest.c from the reference, linked against our library."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "vtest224_b200")):
        pytest.skip("drop-in binaries not built (reference checkout absent)")
    for exe in ("vtest224_b200", "vdecode_b200", "hybridtest_b200"):
        out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(ref, exe)], capture_output=True, text=True, check=True).stdout
        assert "create_viterbi224" in out and "update_viterbi224_blk" in out
        dyn = subprocess.run(["readelf", "-d", os.path.join(ref, exe)], capture_output=True, text=True).stdout
        assert "libviterbi224_b200.so" in dyn


def test_block_driver_is_built_and_fails_loudly_without_a_gpu(built):
    """isee3-decoder_b200/bin/vdecode_block links against the in-tree library (relative rpath) and, on a box without a
    CUDA device, refuses to run instead of falling back to anything: exit code 1 and the library's reason on stderr."""
    exe = os.path.join(ROOT, "isee3-decoder_b200", "bin", "vdecode_block")
    assert os.path.exists(exe)
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libviterbi224_b200.so" in ldd and "not found" not in ldd
    if v224.device_count() == 0:
        r = subprocess.run([exe, "-q"], input=b"\x80" * 64, capture_output=True, timeout=60)
        assert r.returncode == 1 and r.stdout == b""
        assert b"create_viterbi224 failed" in r.stderr and b"no CUDA device" in r.stderr
