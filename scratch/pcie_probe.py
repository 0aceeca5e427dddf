"""PCIe ceiling on this box: H2D alone, D2H alone, both at once (pinned host memory, 1 GiB each)."""
import time, torch
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n / dt / 1e9
run(True, True, 1)
print("H2D alone  %.1f GB/s" % run(True, False))
print("D2H alone  %.1f GB/s" % run(False, True))
print("both       %.1f GB/s per direction" % run(True, True))
# chunked: 64 MiB pieces like the codec's staging
c = 64 << 20
def run_chunked(reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for o in range(0, n, c):
            with torch.cuda.stream(s1): d_a[o:o+c].copy_(h_a[o:o+c], non_blocking=True)
            with torch.cuda.stream(s2): h_b[o:o+c].copy_(d_b[o:o+c], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n / dt / 1e9
print("both, 64 MiB chunks  %.1f GB/s per direction" % run_chunked())
import os; print("cpus", os.cpu_count()); os.system("grep MemTotal /proc/meminfo; nvidia-smi topo -m 2>/dev/null | head -5; lscpu | grep -E 'Model name|Socket|NUMA' ")
# mixed sizes like the codec: H2D in 64 MiB chunks + D2H in 1.6 MB pieces (encode), and the mirror (decode)
def run_mixed(h2d_c, d2h_c, h2d_total, d2h_total, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        oh = od = 0
        while oh < h2d_total or od < d2h_total:
            if oh < h2d_total:
                e = min(oh + h2d_c, h2d_total)
                with torch.cuda.stream(s1): d_a[oh:e].copy_(h_a[oh:e], non_blocking=True)
                oh = e
            # keep the two directions proportional
            while od < d2h_total and od * h2d_total <= oh * d2h_total:
                e = min(od + d2h_c, d2h_total)
                with torch.cuda.stream(s2): h_b[od:e].copy_(d_b[od:e], non_blocking=True)
                od = e
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * (h2d_total + d2h_total) / dt / 1e9
M = 1 << 20
print("encode-like  H2D 1024 MiB in 64 MiB + D2H 389 MiB in 1.6 MB : %.1f GB/s total" % run_mixed(64*M, 1600000, 1024*M, 389*M))
print("encode-like  D2H pieces 25 MB                                : %.1f GB/s total" % run_mixed(64*M, 25*M, 1024*M, 389*M))
print("decode-like  H2D 389 MiB in 25 MB + D2H 1024 MiB in 64 MiB    : %.1f GB/s total" % run_mixed(25*M, 64*M, 389*M, 1024*M))
print("full duplex equal, H2D 64 MiB chunks + D2H 1.6 MB pieces      : %.1f GB/s total" % run_mixed(64*M, 1600000, 1024*M, 1024*M))
print("full duplex equal, 64 MiB both                                : %.1f GB/s total" % run_mixed(64*M, 64*M, 1024*M, 1024*M))
