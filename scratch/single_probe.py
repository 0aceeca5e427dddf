"""Latency of the single-frame C++ drop-in calls (dbde_pack_frame / dbde_unpack_frame)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("dbce-video-cpp_b200")
d = pkg.DropIn()
rng = np.random.default_rng(1)
for W, H in [(2536, 2048), (2048, 2048), (1001, 1003)]:
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    for it in range(4):
        t0 = time.perf_counter(); rec = d.pack_frame(7, img); t1 = time.perf_counter()
        out = d.unpack_frame(rec, W, H); t2 = time.perf_counter()
        print("%dx%d call %d: pack %.2f ms  unpack %.2f ms" % (W, H, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
