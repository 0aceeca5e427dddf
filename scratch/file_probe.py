"""f-1 at scale: write a 2048x2048 micro video to a .dbde file with dbde_b200_writer_*, read it back with
dbde_b200_reader_* (pinned host buffers, C ABI called directly) and with the reference-compatible
walker; throughput + bit-exactness."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, synth, oracle
pkg = importlib.import_module("dbce-video-cpp_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
d = sys.argv[3] if len(sys.argv) > 3 else "/dev/shm"
W = H = 2048; px = W * H
c = pkg.Codec(0); lib = c.lib
torch.cuda.set_device(0)
t = torch.empty(N * px, dtype=torch.uint8, device="cuda")
synth.gen_frames_device("micro", N, W, H, t.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
hin, hout = c.pinned(N * px), c.pinned(batch * px)
lib.dbde_b200_memcpy_d2h(c.h, hin.ptr, t.data_ptr(), N * px)
path = os.path.join(d, "probe.dbde").encode()

def write():
    w = C.c_void_p()
    assert lib.dbde_b200_writer_open(c.h, path, W, H, 100.0, 0, C.byref(w)) == 0
    for a in range(0, N, batch):
        n = min(batch, N - a)
        assert lib.dbde_b200_writer_append(w, hin.ptr + a * px, n) == 0, lib.dbde_b200_file_last_error()
    nf, nb = C.c_uint64(), C.c_uint64()
    assert lib.dbde_b200_writer_close(w, C.byref(nf), C.byref(nb)) == 0
    return nf.value, nb.value

def read(check):
    r, w_, h_, hz = C.c_void_p(), C.c_int(), C.c_int(), C.c_double()
    assert lib.dbde_b200_reader_open(c.h, path, batch, C.byref(w_), C.byref(h_), C.byref(hz), C.byref(r)) == 0
    idx = np.zeros(batch, dtype=np.uint64); st = np.zeros(batch, dtype=np.uint32); done = 0
    while True:
        n = lib.dbde_b200_reader_next(r, hout.ptr, batch, idx.ctypes.data, st.ctypes.data)
        assert n >= 0
        if n == 0: break
        if check:
            assert not st[:n].any() and idx[:n].tolist() == list(range(done, done + n))
            assert np.array_equal(hout.array[:n * px], hin.array[done * px:(done + n) * px])
        done += n
    lib.dbde_b200_reader_close(r)
    return done

write(); assert read(True) == N
t0 = time.perf_counter(); nf, nb = write(); tw = time.perf_counter() - t0
print("write: %d frames, %.1f MB file in %.3f s = %.0f fps = %.2f GB/s raw" % (nf, nb / 1e6, tw, nf / tw, nf * px / tw / 1e9))
t0 = time.perf_counter(); n = read(False); tr = time.perf_counter() - t0
print("read : %d frames in %.3f s = %.0f fps = %.2f GB/s raw" % (n, tr, n / tr, n * px / tr / 1e9))
ora = oracle.best()
want, sizes = ora.pack_frames(hin.array[:8 * px].reshape(8, H, W), 0)
got = np.fromfile(path.decode(), dtype=np.uint8, count=28 + len(want))
assert (got[28:] == want).all(), "file differs from the reference's records"
dr = pkg.DropIn()
t0 = time.perf_counter(); vh, frames = dr.walk_file(path.decode(), 16); tk = time.perf_counter() - t0
assert len(frames) == N and all(np.array_equal(img.ravel(), hin.array[i * px:(i + 1) * px]) for i, (_, img) in enumerate(frames[:64]))
print("walker (dbde_walk_a_file, one frame per call, 16 buffered, pageable image): %d frames in %.3f s = %.0f fps" % (N, tk, N / tk))
os.unlink(path.decode())
