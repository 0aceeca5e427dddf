import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


def rand_frame(rng, W, H, style):
    """Must stay identical to tests/golden/make_golden.py:rand_frame."""
    if style == "noise":
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if style == "flat":
        return np.full((H, W), rng.integers(0, 256), dtype=np.uint8)
    img = np.zeros((H, W), dtype=np.uint8)
    for ty in range(0, H, 8):
        for tx in range(0, W, 8):
            k = int(rng.integers(0, 9))
            rg = (1 << k) - 1
            mn = int(rng.integers(0, 256 - rg))
            blk = mn + (rng.integers(0, 256, (8, 8)) & rg)
            img[ty:ty + 8, tx:tx + 8] = blk[:min(8, H - ty), :min(8, W - tx)]
    return img


def golden_random_frames(golden):
    """Re-create the golden random frames (same rng order as make_golden.py)."""
    rng = np.random.default_rng(20261018)
    out = []
    for c in golden["random_frames"]["cases"]:
        img = rand_frame(rng, c["W"], c["H"], c["style"])
        out.append((c, img))
    return out


def build_example(name="batched_roundtrip"):
    """Compile examples/<name>.cpp against the in-tree library (g++ only: the example is host C++ over the
    C ABI and the reference-compatible header).  -> path of the executable"""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "dbce-video-cpp_b200")
    exe = os.path.join(tempfile.mkdtemp(prefix="dbde_example_"), name)
    subprocess.run(["g++", "-O2", "-std=c++14", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", name + ".cpp"), "-L" + lib, "-ldbde_b200", "-Wl,-rpath," + lib, "-o", exe],
                   check=True)
    return exe
