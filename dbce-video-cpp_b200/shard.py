"""Frame-range sharding for one-process-per-GPU runs.

Frames are independent, so rank r of R takes the contiguous range [r*N/R, (r+1)*N/R) and the ranks'
streams concatenate in rank order: the only "collective" is an exclusive prefix sum over the R shard
byte counts, done on the host.  These helpers hold that logic so it can be tested on CPU with the
gloo backend (tests/test_shard_gloo.py); the data path itself never communicates between GPUs.
"""
import numpy as np


def frame_range(rank, world, nframes):
    """Contiguous, balanced, order-preserving split -- the same rule as the C ABI's sharded calls."""
    return (nframes * rank) // world, (nframes * (rank + 1)) // world


def concat_shards(shards):
    """shards: list (in rank order) of (stream bytes, offsets[n_r+1]) -> (stream, offsets[N+1])."""
    streams, offs, base = [], [], 0
    for stream, o in shards:
        o = np.asarray(o, dtype=np.uint64)
        n = len(o) - 1
        assert int(o[n]) == len(stream)
        offs.append(o[:n] + np.uint64(base))
        streams.append(np.asarray(stream, dtype=np.uint8))
        base += len(stream)
    offs.append(np.array([base], dtype=np.uint64))
    return np.concatenate(streams) if streams else np.zeros(0, np.uint8), np.concatenate(offs)


def gather_stream(dist, stream, offsets, dst=0):
    """Gather every rank's (stream, offsets) on `dst` with torch.distributed (any backend that moves
    CPU tensors, e.g. gloo) and return the concatenated (stream, offsets) there, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([len(stream), len(offsets)], dtype=torch.int64)
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, mine)
    max_s = max(int(s[0]) for s in sizes)
    max_o = max(int(s[1]) for s in sizes)
    pad_s = torch.zeros(max_s, dtype=torch.uint8)
    pad_s[:len(stream)] = torch.from_numpy(np.ascontiguousarray(stream))
    pad_o = torch.zeros(max_o, dtype=torch.int64)
    pad_o[:len(offsets)] = torch.from_numpy(np.ascontiguousarray(offsets).astype(np.int64))
    gs = [torch.zeros(max_s, dtype=torch.uint8) for _ in range(world)] if rank == dst else None
    go = [torch.zeros(max_o, dtype=torch.int64) for _ in range(world)] if rank == dst else None
    dist.gather(pad_s, gs, dst=dst)
    dist.gather(pad_o, go, dst=dst)
    if rank != dst:
        return None
    shards = [(gs[r][:int(sizes[r][0])].numpy(), go[r][:int(sizes[r][1])].numpy().astype(np.uint64)) for r in range(world)]
    return concat_shards(shards)
