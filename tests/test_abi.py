"""CPU tests of the drop-in boundary: the shared library builds, loads, and exports every symbol
that include/*.h declares (no compute calls -- there is no GPU here), and refuses to run
without a B200 instead of falling back."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("dbce-video-cpp_b200")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    return pkg.load()


def test_c_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "dbde_b200.h")).read()
    declared = set(re.findall(r"\b(dbde_b200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dbde_b200_error"}
    assert declared == set(pkg.C_SYMBOLS), declared ^ set(pkg.C_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_cxx_dropin_symbols_match_reference_mangling(lib):
    """The 15 functions of dbde_util.h must carry the reference's mangled names."""
    hdr = open(os.path.join(ROOT, "include", "dbde_util.h")).read()
    declared = set(re.findall(r"\b(dbde_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg.CXX_SYMBOLS), declared ^ set(pkg.CXX_SYMBOLS)
    for name, mangled in pkg.CXX_SYMBOLS.items():
        assert getattr(lib, mangled) is not None, name


def test_size_helpers_are_pure_host_arithmetic(lib):
    assert lib.dbde_b200_frame_record_bound(2048, 2048) == 32 + 66 * 65536 == 4325408
    assert lib.dbde_b200_frame_record_bound(1001, 1003) == 1047848
    assert lib.dbde_b200_frame_record_bound(10, 10) == 296
    assert lib.dbde_b200_slot_stride(10, 10) == 304 and lib.dbde_b200_slot_stride(2048, 2048) == 4325408
    assert lib.dbde_b200_stream_bound(10, 10, 3) == 3 * 304 + 16


def test_index_stream_host_pointer_chase(lib):
    """next = cur + 32 + 2wh + 8*n64 (dbde_util.cpp:301,327); pure host code."""
    import numpy as np
    import oracle
    import synth
    fr = synth.gen_frames("mix", 3, 40, 24)
    stream, sizes = oracle.port.pack_frames(fr, 0)
    offs = np.zeros(8, dtype=np.uint64)
    n = lib.dbde_b200_index_stream(stream.ctypes.data, stream.nbytes, 40, 24, offs.ctypes.data, 7)
    assert n == 3
    assert offs[:4].tolist() == [0] + np.cumsum(sizes).tolist()
    # a torn last record is not indexed
    n = lib.dbde_b200_index_stream(stream.ctypes.data, stream.nbytes - 1, 40, 24, offs.ctypes.data, 7)
    assert n == 2


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.dbde_b200_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.dbde_b200_last_error() or b"CUDA" in lib.dbde_b200_last_error()
    with pytest.raises(pkg.DbdeError):
        pkg.Codec(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dbce-video-cpp_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "oracle/" not in src.replace("never touches oracle/", ""), fn


def test_video_header_hz_as_integer_variant_is_host_only(lib):
    """DBDE_HZ_AS_INTEGER (dbde_util.cpp:203-204,352-353) through the C++ drop-in symbols: pure host
    marshalling, so it runs without a GPU.  Checked against the oracle's variant."""
    import numpy as np
    import oracle
    d = pkg.DropIn()
    try:
        pkg.set_format_variants(False, True)
        assert pkg.get_format_variants() == (False, True)
        got = d.pack_video_header(3, 480, 640, 29.97)
        want = oracle.best_variants().pack_video_header(3, 480, 640, 29.97)
        assert (np.asarray(got) == want).all() and got[20:].tolist() == [30, 0, 0, 0, 0, 0, 0, 0]
        assert d.unpack_video_header(got) == (28, (3, 480, 640, 30.0))
    finally:
        pkg.set_format_variants(False, False)
    assert d.pack_video_header(3, 480, 640, 29.97).tobytes() == oracle.port.pack_video_header(3, 480, 640, 29.97).tobytes()


def test_integration_example_builds_and_refuses_to_run_without_a_gpu(lib):
    """examples/batched_roundtrip.cpp is INTEGRATION.md's batched binding, complete: it must compile and
    link against the two public headers alone, and without a B200 it must stop with the library's
    message rather than produce a file through some other path"""
    import subprocess
    import tempfile
    from conftest import build_example
    exe = build_example()
    out = os.path.join(tempfile.mkdtemp(), "x.dbde")
    r = subprocess.run([exe, out, "64", "64", "2"], capture_output=True, text=True)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_indexers_are_pure_host_arithmetic(lib):
    """dbde_b200_index_stream / _index_stream16: the pointer chase over n64 the reference's walker does
    implicitly (dbde_util.cpp:301,327), on streams made by the oracle -- no GPU involved"""
    import numpy as np
    import oracle
    import synth
    W, H, N = 77, 45, 9
    fr = synth.gen_frames("mix", N, W, H)
    for stream, sizes, fn in ((oracle.best().pack_frames(fr, 0)) + (lib.dbde_b200_index_stream,),
                              (oracle.port16.pack_frames(fr.astype(np.uint16) * 200, 0)) + (lib.dbde_b200_index_stream16,)):
        offs = np.zeros(N + 5, dtype=np.uint64)
        buf = np.concatenate([stream, stream[:40]])              # a torn tenth record must not be counted
        n = fn(buf.ctypes.data, buf.nbytes, W, H, offs.ctypes.data, N + 4)
        assert n == N and offs[:N + 1].tolist() == [0] + np.cumsum(sizes).tolist()
        assert fn(buf.ctypes.data, buf.nbytes, W, H, offs.ctypes.data, 4) == 4      # max_frames is honoured
        assert fn(buf.ctypes.data, 10, W, H, offs.ctypes.data, 4) == 0


def test_library_exports_everything_the_reference_object_exports(lib):
    """nm of libdbde_b200.so must contain every function symbol of the reference's dbde_util.o -- including
    dbde_advance_file_buffer (dbde_util.cpp:394), which the reference exports without declaring it in
    dbde_util.h -- so anything that links against the reference object links against this library."""
    import subprocess
    ref_o = os.path.join(ROOT, "oracle", "_ref", "dbde_util.o")
    if os.path.exists(ref_o):
        out = subprocess.run(["nm", "--defined-only", ref_o], capture_output=True, text=True, check=True).stdout
        want = {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] == "T"}
        assert len(want) == 16, want
    else:                                   # the GPU box: the reference object is not there, its symbol list is
        want = set(pkg.CXX_SYMBOLS.values()) | {"_Z24dbde_advance_file_bufferR16dbde_file_walker"}
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    have = {l.split()[2] for l in out.splitlines() if len(l.split()) == 3}
    assert want <= have, want - have


_GUARD_PAGE_PROBE = r"""
import ctypes as C, importlib, mmap, sys
sys.path.insert(0, %r)
pkg = importlib.import_module("dbce-video-cpp_b200")
d = pkg.DropIn()
libc = C.CDLL(None, use_errno=True)
page = mmap.PAGESIZE
m = mmap.mmap(-1, 2 * page)
base = C.addressof(C.c_char.from_buffer(m))
assert libc.mprotect(C.c_void_p(base + page), C.c_size_t(page), 0) == 0       # the second page faults on any access
W = H = 64; wh = 64
# (1) frame record: 20-byte header + nb, with nb wrong, ending exactly at the guard page
rec = base + page - 24
C.memmove(rec, (2).to_bytes(4, "little") + (7).to_bytes(8, "little") + bytes(8) + (wh + 1).to_bytes(4, "little"), 24)
img = (C.c_uint8 * (W * H))(*([0xCD] * (W * H)))
p = C.cast(rec, C.POINTER(C.c_uint8))
fh = d._unpack_frame(C.byref(p), W, H, img)
assert fh.u64s == 0xFFFFFFFF and fh.index == 7 and C.addressof(p.contents) == rec + 20
# (2) image block: nb right, nm wrong, ending at the guard page
blk = base + page - (8 + wh)
C.memmove(blk, wh.to_bytes(4, "little") + bytes([3] * wh) + (wh - 1).to_bytes(4, "little"), 8 + wh)
assert d._unpack_image(C.cast(blk, C.POINTER(C.c_uint8)), W, H, img) == 0
assert bytes(img) == bytes([0xCD] * (W * H))
print("ok")
"""


def test_malformed_blocks_are_rejected_without_reading_past_what_the_reference_reads(lib):
    """dbde_unpack_image returns 0 right after reading nb when nb != w*h, and after nm when nm != w*h
    (dbde_util.cpp:295-299): the drop-in must not touch a byte beyond those fields either.  The fields sit
    at the very end of a mapped page followed by a PROT_NONE page; an over-read is a segfault.  Host code
    only (the check precedes any GPU work), so this runs without a GPU; in a subprocess because the failure
    mode is a crash."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", _GUARD_PAGE_PROBE % ROOT], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout[-500:], r.stderr[-1500:])
