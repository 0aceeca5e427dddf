"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU code
through oracle/_ref, or the oracle port) prints one JSON line with the keys the driver reads, and
our arm refuses to run without a CUDA device (there is no CPU path to fall back to)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "4", "--width", "256", "--height", "256"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "GB/s"
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["dtype"] == "u8"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "4", "--width", "64", "--height", "64"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return          # on a GPU box this is covered by the gpu tests / the bench itself
    r = _run(["--steps", "1", "--warmup", "1", "--frames", "2", "--width", "64", "--height", "64", "--no-e2e"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr) or "no CPU" in (r.stdout + r.stderr)
