#!/usr/bin/env python
"""DBDE encode+decode benchmark on B200 (the contract in the task statement, tier framing (4)).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                          the reference's CPU path

Workload (BASELINE.json configs[1]): 2048x2048 U8 synthetic microscopy-like video, 1000 frames per
GPU; one STEP = encode the whole batch, then decode it again (device-resident).  Metric: raw-pixel
GB/s = pixels encoded + pixels decoded per second (2 * frames * W * H / step time), summed over
GPUs; frames/s for each direction are reported next to it.  Frames are independent, so N GPUs
take N disjoint frame ranges (weak scaling: 1000 frames each) with no data-path collective;
torch.distributed (NCCL) is used only for the timing barrier and the max over ranks.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dbde_encode_decode_raw_pixel_throughput"
UNIT = "GB/s"
W = H = 2048
KIND = "micro"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-frames", type=int, default=0, help="frames per GPU in the end-to-end leg (0 = all that can be pinned)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kind", default=KIND, choices=["micro", "mix", "low", "noise"])
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--height", type=int, default=H)
    return ap.parse_args()


def workload_config(a, extra=None):
    cfg = {"workload": "%dx%d U8 synthetic '%s' video (SURVEY 8d, seed 42), %d frames per GPU, encode+decode per step"
                       % (a.width, a.height, a.kind, a.frames),
           "frames_per_gpu": a.frames, "width": a.width, "height": a.height, "generator": a.kind,
           "sharding": "contiguous frame ranges per GPU, no collective",
           "l2": "inputs (%.1f GB per direction) exceed the 126 MB L2; no explicit flush"
                 % (a.frames * a.width * a.height / 1e9)}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML as fast as the calls return (~1 ms; power
    every 8th sample) on a thread -- the timed region is tens to hundreds of milliseconds, which
    nvidia-smi -lms cannot resolve.  stop(t0, t1) keeps only the samples taken inside the timed
    region [t0, t1] (time.perf_counter)."""

    def __init__(self, index):
        import threading
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        power, k = 0.0, 0
        while not self._stop.is_set():
            try:
                if k % 8 == 0:
                    power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                k += 1
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), power,
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.0005)

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, b in bits.items() if any(s[3] & b for s in inside))
        if inside:
            out.update(sm_mhz=float(np.median([s[1] for s in inside])), sm_max_mhz=float(self.max_mhz), reasons=reasons,
                       samples=len(inside), power_w_max=max(s[2] for s in inside))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_rates(frames, threads, target_s=1.0):
    """Time the reference's own CPU implementation (oracle/_ref: dbde_util.cpp compiled unmodified)
    on `threads` host threads over contiguous frame ranges.  -> dict of rates."""
    import oracle
    if oracle.ref is None:
        raise RuntimeError("oracle/_ref/libdbde_ref.so missing: build it where /root/reference exists")
    n, Hh, Ww = frames.shape
    # calibrate repetitions so each direction runs for about target_s
    s, slots, sizes = oracle.ref_encode_mt(frames, threads, 1)
    reps_e = max(1, int(target_s / max(s, 1e-4)))
    s_e, slots, sizes = oracle.ref_encode_mt(frames, threads, reps_e)
    s, dec, bad = oracle.ref_decode_mt(slots, Ww, Hh, threads, 1)
    reps_d = max(1, int(target_s / max(s, 1e-4)))
    s_d, dec, bad = oracle.ref_decode_mt(slots, Ww, Hh, threads, reps_d)
    assert bad == 0 and (dec == frames).all(), "reference round trip failed"
    enc_fps, dec_fps = n * reps_e / s_e, n * reps_d / s_d
    px = Ww * Hh
    t_pair = 1.0 / enc_fps + 1.0 / dec_fps            # seconds to encode AND decode one frame
    return {"encode_fps": enc_fps, "decode_fps": dec_fps, "value": 2 * px / t_pair / 1e9,
            "cpu_seconds": threads * (s_e + s_d), "reps": [reps_e, reps_d], "sizes": sizes}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import synth
    threads = os.cpu_count() or 1
    nsample = min(a.frames, 64)
    frames = synth.gen_frames(a.kind, nsample, a.width, a.height)
    per = []
    for _ in range(a.warmup):
        cpu_reference_rates(frames, threads, target_s=0.2)
    t0 = time.time()
    target = max(0.1, min(0.5, 15.0 / max(a.steps, 1)))       # ~30 s of timed CPU work in total, whatever K is
    for _ in range(a.steps):
        per.append(cpu_reference_rates(frames, threads, target_s=target))
    wall = time.time() - t0
    val = float(np.mean([p["value"] for p in per]))
    enc = float(np.mean([p["encode_fps"] for p in per]))
    dec = float(np.mean([p["decode_fps"] for p in per]))
    sample = "%d frames of the same workload, dbde_pack_frame then dbde_unpack_frame, ~%.2f s per direction per step" % (nsample, target)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * wall / max(a.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(a), "encode_fps": enc, "decode_fps": dec,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "unmodified dbde_util.cpp (g++ -O3 -march=corei7, SSE4.1) via oracle/_ref, %d std::threads over "
                    "contiguous frame ranges; host only, no GPU" % threads}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import synth
    pkg = importlib.import_module("dbce-video-cpp_b200")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the DBDE B200 codec has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    codec = pkg.Codec(local)
    Ww, Hh, N = a.width, a.height, a.frames
    px = Ww * Hh
    wh = ((Ww + 7) // 8) * ((Hh + 7) // 8)
    f0 = rank * N                                    # this GPU's contiguous frame range

    # ---- resident buffers (torch owns the device memory; the codec gets raw pointers)
    cap = codec.stream_bound(Ww, Hh, N)
    frames = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
    stream_buf = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    decoded = torch.empty(N * px + 64, dtype=torch.uint8, device=dev)
    offs = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    sizes = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    status = torch.zeros(N, dtype=torch.int32, device=dev)
    delta = (16 - (32 + 2 * wh) % 16) % 16           # puts every frame's U64 words on 8/16-byte boundaries
    out_ptr = stream_buf.data_ptr() + delta
    cs = torch.cuda.current_stream().cuda_stream
    synth.gen_frames_device(a.kind, N, Ww, Hh, frames.data_ptr(), seed=42, f0=f0, stream=cs)
    torch.cuda.synchronize()

    def encode():
        codec.encode_device(frames.data_ptr(), Ww, Hh, f0, N, out_ptr, cap, offs.data_ptr(), sizes.data_ptr(), cs)

    def decode(total):
        # the decoder reads each record from its slot (offs[i] = i * slot_stride)
        codec.decode_device(out_ptr, cap, offs.data_ptr(), Ww, Hh, N, decoded.data_ptr(), status.data_ptr(), None, cs)

    # ---- correctness gate before any timing: round trip + a sample against the oracle
    encode()
    torch.cuda.synchronize()
    total = int(sizes[:N].sum().item())           # bytes of all records (what a file would hold)
    decode(total)
    torch.cuda.synchronize()
    assert int(status.abs().sum().item()) == 0, "decode rejected frames"
    assert torch.equal(frames[:N * px], decoded[:N * px]), "decode(encode(x)) != x"
    n64_total = (total - N * (32 + 2 * wh)) // 8
    alg_bytes = N * px + 2 * N * wh + 8 * n64_total   # per launch, same both directions

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps
    for _ in range(a.warmup):
        encode(); decode(total)
    barrier()
    launches0 = codec.launches()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    start.record()
    for k in range(a.steps):
        ev[k][0].record()
        encode()
        ev[k][1].record()
        decode(total)
        ev[k][2].record()
    end.record()
    barrier()
    t_host1 = time.perf_counter()
    elapsed_ms = start.elapsed_time(end)
    launches = codec.launches() - launches0
    clocks = sampler.stop(t_host0, t_host1) if sampler else None
    enc_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    dec_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    t = torch.tensor([elapsed_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, enc_ms_max, dec_ms_max = [float(x) for x in t.tolist()]
    ms_per_step = elapsed_ms / a.steps
    value = world * 2 * N * px / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region.
    # A step is still "encode N frames + decode N frames", but the two directions run CONCURRENTLY on
    # two contexts driven by two host threads (the C ABI's contract: one context per thread), so the
    # encoder's H2D of raw frames overlaps the decoder's D2H of raw frames on the full-duplex PCIe
    # link.  Step k decodes the stream step k-1 encoded (same bytes every step).  The strictly
    # sequential figure (encode_host, then decode_host, one context) is reported next to it.
    e2e = None
    if not a.no_e2e:
        from concurrent.futures import ThreadPoolExecutor
        codec2 = pkg.Codec(local)
        # pinned host memory: frames + decoded frames + two worst-case stream buffers ~ 4.1 x raw bytes.
        # Use the whole N-frame batch when the box can pin it for every rank, else the largest prefix
        # that fits in a third of MemAvailable (throughput is steady-state either way).
        Ne = N
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
            per_frame = 2 * px + 2 * codec.stream_bound(Ww, Hh, 1)
            Ne = max(16, min(N, int(avail / 3 / max(world, 1) / per_frame)))
            if a.e2e_frames > 0:
                Ne = min(N, a.e2e_frames)
        except Exception:
            pass
        N_full, N = N, Ne                              # the e2e leg below runs on the first Ne frames
        cap = codec.stream_bound(Ww, Hh, N)
        total = int(sizes[:N].sum().item())
        h_frames = codec.pinned(N * px)
        h_streams = [codec.pinned(cap), codec.pinned(cap)]
        h_dec = codec.pinned(N * px)
        h_offs = [np.zeros(N + 1, dtype=np.uint64), np.zeros(N + 1, dtype=np.uint64)]
        h_status = np.zeros(N, dtype=np.uint32)
        torch.cuda.synchronize()
        codec.lib.dbde_b200_memcpy_d2h(codec.h, h_frames.ptr, frames.data_ptr(), N * px)

        def enc_host(k):
            codec.encode_host_raw(h_frames.ptr, Ww, Hh, f0, N, h_streams[k % 2].ptr, cap, h_offs[k % 2].ctypes.data)

        def dec_host(k, c=None):
            (c or codec2).decode_host_raw(h_streams[k % 2].ptr, int(h_offs[k % 2][N]), h_offs[k % 2].ctypes.data, Ww, Hh, N,
                                          h_dec.ptr, h_status.ctypes.data, None)

        # warm-up (allocates the staging slots of both contexts) + correctness of the host path
        enc_host(0); enc_host(1); dec_host(1); dec_host(0, codec)
        assert int(h_offs[0][N]) == total and int(h_offs[1][N]) == total and not h_status.any()
        assert np.array_equal(h_dec.array, h_frames.array), "e2e round trip differs"
        # sequential: one context, encode then decode
        barrier()
        t0 = time.perf_counter()
        for k in range(a.e2e_steps):
            enc_host(k); dec_host(k, codec)
        dt_seq = time.perf_counter() - t0
        # concurrent: encode(k) || decode(k-1)
        pool = ThreadPoolExecutor(2)
        barrier()
        t0 = time.perf_counter()
        for k in range(a.e2e_steps):
            fe, fd = pool.submit(enc_host, k), pool.submit(dec_host, k + 1)
            fe.result(); fd.result()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        pool.shutdown()
        assert not h_status.any() and np.array_equal(h_dec.array, h_frames.array), "e2e round trip differs"
        te = torch.tensor([dt, dt_seq], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt, dt_seq = [float(x) for x in te.tolist()]
        e2e = {"value": world * 2 * N * px * a.e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(N * px + total + 8 * N), "d2h_bytes_per_step": int(total + 8 * (N + 1) + N * px + 4 * N),
               "steps": a.e2e_steps, "ms_per_step": 1000 * dt / a.e2e_steps, "frames_per_gpu_per_step": int(N),
               "sequential": {"value": world * 2 * N * px * a.e2e_steps / dt_seq / 1e9, "ms_per_step": 1000 * dt_seq / a.e2e_steps},
               "path": "dbde_b200_encode_host || dbde_b200_decode_host on pinned host buffers: two contexts on two host "
                       "threads, each chunked through 3 device staging slots; step k decodes the stream of step k-1"}
        for b in [h_frames, h_dec] + h_streams:
            b.free()
        codec2.close()
        N = N_full
        cap = codec.stream_bound(Ww, Hh, N)
        total = int(sizes[:N].sum().item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (algorithmic bytes / CUDA-event duration)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    enc_gbs, dec_gbs = alg_bytes / (enc_ms * 1e-3) / 1e9, alg_bytes / (dec_ms * 1e-3) / 1e9
    dom = "encode" if enc_ms >= dec_ms else "decode"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))      # ncu dram__bytes_read+write per launch, captured for ONE workload
            if tj.get("workload") == {"kind": a.kind, "width": Ww, "height": Hh, "frames": N}:
                traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "dbde_%s_kernel" % dom, "achieved": enc_gbs if dom == "encode" else dec_gbs,
                "peak": peak, "unit": "GB/s", "frac": (enc_gbs if dom == "encode" else dec_gbs) / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                "encode": {"ms": enc_ms, "GBps": enc_gbs, "frac": enc_gbs / peak},
                "decode": {"ms": dec_ms, "GBps": dec_gbs, "frac": dec_gbs / peak,
                           "note": "scan pre-pass + unpack kernel, timed together"}}

    # ---- the reference's CPU path on this box's host cores, bounded sample
    cpu = None
    if not a.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            ns = min(N, 64)
            sample = frames[:ns * px].cpu().numpy().reshape(ns, Hh, Ww)
            r = cpu_reference_rates(sample, threads, target_s=1.0)
            cpu = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": "first %d frames of this rank's batch, dbde_pack_frame/dbde_unpack_frame via oracle/_ref, "
                             "~1 s per direction (%.0f core-seconds)" % (ns, r["cpu_seconds"]),
                   "encode_fps": r["encode_fps"], "decode_fps": r["decode_fps"]}
        except Exception as ex:            # the baseline must never take the bench line down
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % ex}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(a, {"n_gpus": world}),
            "encode_fps": world * N / (enc_ms_max * 1e-3), "decode_fps": world * N / (dec_ms_max * 1e-3),
            "encode_raw_GBps": world * N * px / (enc_ms_max * 1e-3) / 1e9,
            "decode_raw_GBps": world * N * px / (dec_ms_max * 1e-3) / 1e9,
            "compressed_ratio": total / float(N * px), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_note": "per step: 1 encode kernel + 2 decode kernels (scan, unpack); memset excluded",
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    codec.close()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))
