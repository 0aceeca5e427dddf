#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02q.log 2>&1; tail -3 $O/pytest_gpu_r02q.log
timeout 120 python scratch/enc16_only.py 2>&1 | tail -6
timeout 100 python - <<'PY'
import importlib, time, numpy as np, sys
sys.path.insert(0, ".")
pkg = importlib.import_module("dbce-video-cpp_b200")
c = pkg.Codec(0)
rng = np.random.default_rng(0)
N, W, H = 96, 2048, 2048
fr = (1000 + rng.integers(0, 50, (N, H, W))).astype(np.uint16)
for rep in range(2):
    t0 = time.time(); s, offs = c.encode16_host(fr, 0); t1 = time.time()
    dec, st, _ = c.decode16_host(s, offs[:N], W, H); t2 = time.time()
    print("DBDE16 host path, pageable numpy buffers, %d frames of 2048^2: encode %.0f fps (%.1f GB/s raw), decode %.0f fps, ok=%s" % (N, N/(t1-t0), N*W*H*2/(t1-t0)/1e9, N/(t2-t1), bool((dec==fr).all() and (st==0).all())))
PY
