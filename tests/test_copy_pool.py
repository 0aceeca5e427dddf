"""The copy pool behind the pageable-memory path (dbde_copy_pool.h), stress-tested without a GPU: caller threads
relay buffers through bounce memory with the same job / flag / helping-wait pattern as relay_h2d and relay_d2h,
with and without pool threads, more callers than cores included."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stress_binary(tmp_path_factory):
    out = tmp_path_factory.mktemp("cps") / "copy_pool_stress"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "cpp", "copy_pool_stress.cpp"), "-o", str(out)],
                   check=True)
    return str(out)


@pytest.mark.parametrize("workers,threads", [(0, 1), (0, 8), (0, 24), (3, 1), (3, 24), (6, 16)])
def test_copy_pool_relays_under_load(stress_binary, workers, threads):
    env = dict(os.environ, DBDE_B200_COPY_THREADS=str(workers))
    r = subprocess.run([stress_binary, str(threads), "40"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
