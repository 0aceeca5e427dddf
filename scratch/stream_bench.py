"""BASELINE config 5: a long 4096x4096 stream sharded by contiguous frame ranges over all GPUs of the
box, from pinned host memory, through dbde_b200_encode_host_sharded / dbde_b200_decode_host_sharded
(one process, one context + one host thread per GPU, no collective).  The stream is fed through a
bounded ring of pinned batches (168 GB of raw frames do not have to exist at once).

    python scratch/stream_bench.py [total_frames=10000] [batch=256] [kind=micro] [W=4096] [H=4096]
"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, synth, oracle
pkg = importlib.import_module("dbce-video-cpp_b200")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kind = sys.argv[3] if len(sys.argv) > 3 else "micro"
W = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
H = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
G = torch.cuda.device_count()
px = W * H
codecs = [pkg.Codec(g) for g in range(G)]
lib = codecs[0].lib
arr = (C.c_void_p * G)(*[c.h.value for c in codecs])
cap = codecs[0].stream_bound(W, H, B)
# ring of two pinned input batches (distinct content), one pinned stream buffer, one pinned output batch
ring = [codecs[0].pinned(B * px) for _ in range(2)]
h_stream, h_out = codecs[0].pinned(cap), codecs[0].pinned(B * px)
offs = np.zeros(B + 1, dtype=np.uint64); status = np.zeros(B, dtype=np.uint32); index = np.zeros(B, dtype=np.uint64)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)          # creating the contexts left the last GPU current
tmp = torch.empty(B * px + 64, dtype=torch.uint8, device=dev)
for r, buf in enumerate(ring):
    synth.gen_frames_device(kind, B, W, H, tmp.data_ptr(), seed=42, f0=r * B, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    lib.dbde_b200_memcpy_d2h(codecs[0].h, buf.ptr, tmp.data_ptr(), B * px)
del tmp

def enc(buf, first):
    rc = lib.dbde_b200_encode_host_sharded(arr, G, buf.ptr, W, H, first, B, h_stream.ptr, cap, offs.ctypes.data)
    assert rc == 0, lib.dbde_b200_last_error()
def dec():
    rc = lib.dbde_b200_decode_host_sharded(arr, G, h_stream.ptr, int(offs[B]), offs.ctypes.data, W, H, B, h_out.ptr,
                                           status.ctypes.data, index.ctypes.data)
    assert rc == 0, lib.dbde_b200_last_error()

# parity of the sharded stream: first frames of batch 0 against the oracle, and the round trip
enc(ring[0], 0); dec()
assert not status.any() and np.array_equal(h_out.array, ring[0].array) and index.tolist() == list(range(B))
ora = oracle.best()
nchk = min(B, 4)
want, sizes = ora.pack_frames(ring[0].array[:nchk * px].reshape(nchk, H, W), 0)
assert np.array_equal(h_stream.array[:len(want)], want), "sharded stream differs from the oracle"
last = ring[0].array[(B - 1) * px:].reshape(1, H, W)
want, _ = ora.pack_frames(last, B - 1)
assert np.array_equal(h_stream.array[int(offs[B - 1]):int(offs[B])], want), "last record of the sharded stream differs"
nb = (F + B - 1) // B
t0 = time.perf_counter(); comp = 0
for b in range(nb):
    enc(ring[b % 2], b * B); comp += int(offs[B])
te = time.perf_counter() - t0
t0 = time.perf_counter()
for b in range(nb):
    dec()
td = time.perf_counter() - t0
assert not status.any() and np.array_equal(h_out.array, ring[(nb - 1) % 2].array)
n = nb * B
print("%d GPUs, %d frames of %dx%d '%s' in batches of %d (ratio %.3f): encode %.2f s = %.0f fps = %.1f GB/s raw;  "
      "decode %.2f s = %.0f fps = %.1f GB/s raw" % (G, n, W, H, kind, B, comp / (n * px), te, n / te, n * px / te / 1e9,
                                                   td, n / td, n * px / td / 1e9))
