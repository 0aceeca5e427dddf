"""profiling helper: time encode/decode kernels alone on device-resident micro frames"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, synth
pkg = importlib.import_module("dbce-video-cpp_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "micro"
W = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
H = int(sys.argv[5]) if len(sys.argv) > 5 else 2048
check = os.environ.get("CHECK", "1") == "1"
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
c = pkg.Codec(0)
px = W * H; wh = ((W + 7) // 8) * ((H + 7) // 8)
cap = c.stream_bound(W, H, N)
fr = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
out = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
dec = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
offs = torch.zeros(N + 1, dtype=torch.int64, device=dev)
szs = torch.zeros(N + 1, dtype=torch.int64, device=dev)
st = torch.zeros(N, dtype=torch.int32, device=dev)
cs = torch.cuda.current_stream().cuda_stream
synth.gen_frames_device(kind, N, W, H, fr.data_ptr(), stream=cs)
delta = (16 - (32 + 2 * wh) % 16) % 16
op = out.data_ptr() + delta
torch.cuda.synchronize()
def enc(): c.encode_device(fr.data_ptr(), W, H, 0, N, op, cap, offs.data_ptr(), szs.data_ptr(), cs)
enc(); torch.cuda.synchronize(); total = int(szs[:N].sum().item())
def decf(): c.decode_device(op, cap, offs.data_ptr(), W, H, N, dec.data_ptr(), st.data_ptr(), None, cs)
if check:
    decf(); torch.cuda.synchronize()
    print("roundtrip ok:", bool(torch.equal(fr[:N*px], dec[:N*px])), "status", int(st.abs().sum()))
alg = N * px + 2 * N * wh + (total - N * (32 + 2 * wh))
for name, fn in (("encode", enc), ("decode", decf)):
    if name == "decode" and not check: continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s: %.3f ms per %d frames  -> %.0f GB/s algorithmic, %.0f fps" % (name, ms, N, alg / ms / 1e6, N / ms * 1e3))
