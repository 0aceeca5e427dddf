"""Generate tests/golden/golden16.json: fixtures that freeze the DBDE16 extension.

DBDE16 has no reference implementation, so these vectors come from the oracle's own definition
(oracle/dbde_oracle.c, "DBDE16") at the moment the extension was introduced; they exist so that later
changes to the oracle or the kernels cannot silently change the format.  What ties the definition to the
reference is the embedding test in tests/test_oracle.py, not this file.
      python tests/golden/make_golden16.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402


def frames16(case):
    """Deterministic 16-bit frames from integer arithmetic only (no RNG library dependence)."""
    W, H, N, kind = case["W"], case["H"], case["N"], case["kind"]
    f, y, x = np.meshgrid(np.arange(N, dtype=np.uint64), np.arange(H, dtype=np.uint64), np.arange(W, dtype=np.uint64), indexing="ij")
    with np.errstate(over="ignore"):
        h = (x * np.uint64(0x9E3779B97F4A7C15)) ^ (y * np.uint64(0xBF58476D1CE4E5B9)) ^ ((f + np.uint64(1)) * np.uint64(0x94D049BB133111EB))
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xD6E8FEB86659FD93)
        h ^= h >> np.uint64(32)
    if kind == "noise":
        v = h & np.uint64(0xFFFF)
    elif kind == "sensor12":
        v = np.uint64(600) + x // np.uint64(16) + y // np.uint64(8) + (h & np.uint64(31))
    else:                                   # "classes": tile (ty, tx) gets depth (tx + 3 ty + f) % 17
        k = ((x // np.uint64(8)) + np.uint64(3) * (y // np.uint64(8)) + f) % np.uint64(17)
        rg = (np.uint64(1) << k) - np.uint64(1)
        v = np.uint64(1000) + (h & rg)
        v = np.minimum(v, np.uint64(65535))
    return v.astype(np.uint16)


CASES = [dict(W=8, H=8, N=1, kind="classes"), dict(W=10, H=10, N=2, kind="sensor12"), dict(W=17, H=23, N=3, kind="classes"),
         dict(W=64, H=40, N=2, kind="noise"), dict(W=136, H=72, N=4, kind="classes"), dict(W=1001, H=24, N=2, kind="sensor12")]

if __name__ == "__main__":
    out = {"note": "DBDE16 fixtures frozen from the oracle's definition; see make_golden16.py", "cases": []}
    for c in CASES:
        fr = frames16(c)
        stream, sizes = oracle.port16.pack_frames(fr, 11)
        e = dict(c)
        e["sizes"] = [int(s) for s in sizes]
        e["frames_sha256"] = hashlib.sha256(fr.tobytes()).hexdigest()
        e["stream_sha256"] = hashlib.sha256(stream.tobytes()).hexdigest()
        if len(stream) <= 400:
            e["stream_hex"] = stream.tobytes().hex()
        out["cases"].append(e)
    path = os.path.join(os.path.dirname(__file__), "golden16.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, [c["sizes"] for c in out["cases"]])
