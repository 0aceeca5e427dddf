"""CPU test of the N>1 host logic: two processes over gloo each take a contiguous frame range,
"encode" it (with the oracle -- there is no GPU here; the shard/concat logic is what is under
test), and rank 0 reassembles a stream that must equal the single-process stream byte for byte."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import synth

shard = importlib.import_module("dbce-video-cpp_b200.shard")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, W, H, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard.frame_range(rank, world, nframes)
    frames = synth.gen_frames("mix", b - a, W, H, seed=42, f0=a)          # this rank's frames only
    stream, sizes = oracle.port.pack_frames(frames, a)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    out = shard.gather_stream(dist, stream, offs, dst=0)
    # timing protocol of bench.py: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((out[0].tobytes(), out[1].tolist(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_frame_range_sharding_matches_single_process():
    nframes, W, H, world = 7, 40, 24, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    stream, offs, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, sizes = oracle.port.pack_frames(synth.gen_frames("mix", nframes, W, H, seed=42), 0)
    assert stream == want.tobytes()
    assert offs == [0] + np.cumsum(sizes).tolist()
    assert tmax == 2.0


def test_frame_range_is_a_partition():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1000, 10000):
            r = [shard.frame_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_concat_shards_rebases_offsets():
    s0, o0 = np.arange(10, dtype=np.uint8), np.array([0, 4, 10], dtype=np.uint64)
    s1, o1 = np.arange(5, dtype=np.uint8), np.array([0, 5], dtype=np.uint64)
    s, o = shard.concat_shards([(s0, o0), (s1, o1)])
    assert len(s) == 15 and o.tolist() == [0, 4, 10, 15]
