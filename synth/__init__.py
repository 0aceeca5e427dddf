"""Deterministic synthetic video: BASELINE.json's workloads as integer-only
generators, on the CPU (libdbde_gen.so) and on the GPU (libdbde_synth.so, same bytes).
Test / bench DATA only -- there is no codec code here."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = {"noise": 0, "micro": 1, "mix": 2, "low": 3}


def build(verbose=False):
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("synth build failed")


def _lib(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_cpu = None
_gpu = None


def gen_frames(kind, nframes, W, H, seed=42, f0=0):
    """CPU synthetic frames -> (N,H,W) u8."""
    global _cpu
    if _cpu is None:
        _cpu = _lib("libdbde_gen.so")
        _cpu.gen_frames.restype = None
        _cpu.gen_frames.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    out = np.zeros((nframes, H, W), dtype=np.uint8)
    _cpu.gen_frames(KINDS[kind], seed, f0, nframes, W, H, out.ctypes.data)
    return out


def gen_frames_device(kind, nframes, W, H, dev_ptr, seed=42, f0=0, stream=None):
    """Fill nframes*W*H bytes at device pointer `dev_ptr` (asynchronous on `stream`)."""
    global _gpu
    if _gpu is None:
        _gpu = _lib("libdbde_synth.so")
        _gpu.synth_frames_device.restype = C.c_int
        _gpu.synth_frames_device.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p]
    rc = _gpu.synth_frames_device(KINDS[kind], seed, f0, nframes, W, H, dev_ptr, stream)
    if rc:
        raise RuntimeError("synth_frames_device: cuda error %d" % rc)
