"""DBDE16 kernels alone on device-resident 16-bit frames: scratch/enc16_only.py N reps kind W H
kinds: sensor12 (12-bit camera-like: smooth field + 5 bits of noise), noise16 (all depth 16), flat"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("dbce-video-cpp_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kind = sys.argv[3] if len(sys.argv) > 3 else "sensor12"
W = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
H = int(sys.argv[5]) if len(sys.argv) > 5 else 2048
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
c = pkg.Codec(0); lib = c.lib
px = W * H; wh = ((W + 7) // 8) * ((H + 7) // 8)
g = torch.Generator(device=dev); g.manual_seed(42)
if kind == "noise16":
    fr = torch.randint(0, 65536, (N, H, W), device=dev, generator=g, dtype=torch.int32)
elif kind == "flat":
    fr = torch.full((N, H, W), 1234, device=dev, dtype=torch.int32)
else:
    y = torch.arange(H, device=dev).view(1, H, 1); x = torch.arange(W, device=dev).view(1, 1, W)
    fr = (600 + (x // 16) + (y // 8) + torch.randint(0, 32, (N, H, W), device=dev, generator=g, dtype=torch.int32)).to(torch.int32)
fr = fr.to(torch.uint16).contiguous()
stride = int(lib.dbde_b200_slot_stride16(W, H)); cap = stride * N
out = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
dec = torch.zeros_like(fr)
offs = torch.zeros(N + 1, dtype=torch.int64, device=dev); szs = torch.zeros(N + 1, dtype=torch.int64, device=dev)
st = torch.zeros(N, dtype=torch.int32, device=dev)
cs = torch.cuda.current_stream().cuda_stream
def enc(): assert lib.dbde_b200_encode16_device(c.h, fr.data_ptr(), W, H, 0, N, out.data_ptr(), cap, 0, offs.data_ptr(), szs.data_ptr(), cs) == 0
def decf(): assert lib.dbde_b200_decode16_device(c.h, out.data_ptr(), cap, offs.data_ptr(), W, H, N, dec.data_ptr(), st.data_ptr(), None, cs) == 0
enc(); decf(); torch.cuda.synchronize()
total = int(szs[:N].sum().item())
print("roundtrip ok:", bool(torch.equal(fr.view(torch.int16), dec.view(torch.int16))), "status", int(st.abs().sum()), "ratio %.3f" % (total / (2.0 * N * px)))
alg = 2 * N * px + 3 * N * wh + (total - N * (32 + 3 * wh))
for name, fn in (("encode16", enc), ("decode16", decf)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %.3f ms per %d frames  -> %.0f GB/s algorithmic, %.0f fps" % (name, ms, N, alg / ms / 1e6, N / ms * 1e3))
