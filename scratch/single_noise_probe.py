"""why is a ONE-frame noise decode slow?  device-resident timing of 1-frame launches + per-call host timing"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, synth
pkg = importlib.import_module("dbce-video-cpp_b200")
c = pkg.Codec(0)
dev = torch.device("cuda", 0)
for kind in ("micro", "noise"):
    for N in (1, 2, 8):
        W = H = 2048; px = W * H; wh = 65536
        cap = c.stream_bound(W, H, N)
        fr = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
        out = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
        dec = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
        offs = torch.zeros(N + 1, dtype=torch.int64, device=dev); szs = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        st = torch.zeros(N, dtype=torch.int32, device=dev)
        cs = torch.cuda.current_stream().cuda_stream
        synth.gen_frames_device(kind, N, W, H, fr.data_ptr(), stream=cs)
        op = out.data_ptr()
        def enc(): c.encode_device(fr.data_ptr(), W, H, 0, N, op, cap, offs.data_ptr(), szs.data_ptr(), cs)
        def decf(): c.decode_device(op, cap, offs.data_ptr(), W, H, N, dec.data_ptr(), st.data_ptr(), None, cs)
        enc(); decf(); torch.cuda.synchronize()
        assert torch.equal(fr[:N*px], dec[:N*px])
        for name, fn in (("encode", enc), ("decode", decf)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            print("%s N=%d %s: %.1f us per launch" % (kind, N, name, e0.elapsed_time(e1) / 20 * 1e3))
d = pkg.DropIn()
rng = np.random.default_rng(1)
img = rng.integers(0, 256, (2048, 2048), dtype=np.uint8)
rec = d.pack_frame(7, img)
for it in range(8):
    t0 = time.perf_counter(); out = d.unpack_frame(rec, 2048, 2048); t1 = time.perf_counter()
    print("noise unpack_frame call %d: %.3f ms" % (it, (t1 - t0) * 1e3))
