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
