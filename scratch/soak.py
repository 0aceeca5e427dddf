"""Soak: thousands of back-to-back device-resident encode/decode launches at random batch sizes and
geometries; every launch's records and pixels are compared with the first result for that input
(look-back ordering, ticket races and pipeline hand-offs would show up as a mismatch or a hang)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, synth
pkg = importlib.import_module("dbce-video-cpp_b200")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
c = pkg.Codec(0)
cs = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1)
cases = []
for kind, W, H, Nmax in [("micro", 2048, 2048, 64), ("mix", 1001, 1003, 96), ("low", 4096, 4096, 16), ("mix", 264, 40, 512), ("noise", 2048, 512, 64),
                         ("mix", 2049, 9, 64), ("micro", 13, 9, 2000), ("micro", 2304, 2304, 24), ("mix", 1280, 1024, 64),
                         ("mix", 2560, 40, 200)]:
    px = W * H
    fr = torch.empty(Nmax * px + 64, dtype=torch.uint8, device=dev)
    synth.gen_frames_device(kind, Nmax, W, H, fr.data_ptr(), stream=cs)
    cap = c.stream_bound(W, H, Nmax)
    cases.append(dict(kind=kind, W=W, H=H, Nmax=Nmax, px=px, fr=fr, cap=cap, out=torch.empty(cap + 64, dtype=torch.uint8, device=dev),
                      dec=torch.empty(Nmax * px + 64, dtype=torch.uint8, device=dev), offs=torch.zeros(Nmax + 1, dtype=torch.int64, device=dev),
                      szs=torch.zeros(Nmax + 1, dtype=torch.int64, device=dev), st=torch.zeros(Nmax, dtype=torch.int32, device=dev), ref={}))
torch.cuda.synchronize()
t0 = time.time(); it = 0; frames = 0
while time.time() - t0 < secs:
    k = cases[int(rng.integers(0, len(cases)))]
    N = int(rng.integers(1, k["Nmax"] + 1))
    wh = ((k["W"] + 7) // 8) * ((k["H"] + 7) // 8)
    delta = (16 - (32 + 2 * wh) % 16) % 16
    op = k["out"].data_ptr() + delta
    k["out"].zero_()
    c.encode_device(k["fr"].data_ptr(), k["W"], k["H"], 5, N, op, k["cap"], k["offs"].data_ptr(), k["szs"].data_ptr(), cs)
    k["dec"].zero_()
    c.decode_device(op, k["cap"], k["offs"].data_ptr(), k["W"], k["H"], N, k["dec"].data_ptr(), k["st"].data_ptr(), None, cs)
    torch.cuda.synchronize()
    assert int(k["st"][:N].abs().sum()) == 0, (k["kind"], N)
    assert torch.equal(k["fr"][:N * k["px"]], k["dec"][:N * k["px"]]), ("decode", k["kind"], k["W"], N, it)
    sig = (int(k["szs"][:N].sum()), int(k["out"].to(torch.int64).sum()))          # sizes and a checksum of every record byte
    # records of frames 0..N-1 do not depend on N: compare the first frame's record with the first time we saw it
    first = bytes(k["out"][delta:delta + int(k["szs"][0])].cpu().numpy())
    if "first" not in k["ref"]: k["ref"]["first"] = first
    assert k["ref"]["first"] == first, ("record 0 changed", k["kind"], N, it)
    k["ref"].setdefault(N, sig)
    assert k["ref"][N] == sig, ("batch checksum changed", k["kind"], N, it)
    it += 1; frames += N
print("soak ok: %d launches pairs, %d frames in %.1f s, no mismatch, no hang" % (it, frames, time.time() - t0))
