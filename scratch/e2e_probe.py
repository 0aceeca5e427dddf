"""End-to-end host path probe: sequential and concurrent encode/decode for the current
DBDE_B200_SLOTS / DBDE_B200_CHUNK_FRAMES environment."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, synth
from concurrent.futures import ThreadPoolExecutor
pkg = importlib.import_module("dbce-video-cpp_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
W = H = 2048; px = W * H
c1, c2 = pkg.Codec(0), pkg.Codec(0)
cap = c1.stream_bound(W, H, N)
dev = torch.device("cuda", 0)
fr = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
synth.gen_frames_device("micro", N, W, H, fr.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
hf, hd = c1.pinned(N * px), c1.pinned(N * px)
hs = [c1.pinned(cap), c1.pinned(cap)]
ho = [np.zeros(N + 1, dtype=np.uint64), np.zeros(N + 1, dtype=np.uint64)]
st = np.zeros(N, dtype=np.uint32)
c1.lib.dbde_b200_memcpy_d2h(c1.h, hf.ptr, fr.data_ptr(), N * px)
def enc(k): c1.encode_host_raw(hf.ptr, W, H, 0, N, hs[k % 2].ptr, cap, ho[k % 2].ctypes.data)
def dec(k, c=c2): c.decode_host_raw(hs[k % 2].ptr, int(ho[k % 2][N]), ho[k % 2].ctypes.data, W, H, N, hd.ptr, st.ctypes.data, None)
enc(0); enc(1); dec(0); dec(1, c1)
assert np.array_equal(hd.array, hf.array)
R = 3
t0 = time.perf_counter()
for k in range(R): enc(k)
te = (time.perf_counter() - t0) / R
t0 = time.perf_counter()
for k in range(R): dec(k)
td = (time.perf_counter() - t0) / R
pool = ThreadPoolExecutor(2)
t0 = time.perf_counter()
for k in range(R):
    a, b = pool.submit(enc, k), pool.submit(dec, k + 1); a.result(); b.result()
tc = (time.perf_counter() - t0) / R
print("slots=%s chunk=%s : encode %.1f ms  decode %.1f ms  sequential %.1f GB/s  concurrent %.1f ms -> %.1f GB/s" % (
    os.environ.get("DBDE_B200_SLOTS", "3"), os.environ.get("DBDE_B200_CHUNK_FRAMES", "auto"), te * 1e3, td * 1e3,
    2 * N * px / (te + td) / 1e9, tc * 1e3, 2 * N * px / tc / 1e9))
