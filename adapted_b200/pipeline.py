"""Multi-GPU sharding of the hot path: one process per GPU, minibatches dealt round-robin, no data-path collective.

The independent unit of the reference is the minibatch, so rank r of G takes minibatches
r, r+G, r+2G, ... of the read stream; results are fixed-size records that are gathered on the host (rank 0) in
minibatch order -- that is the only communication (torch.distributed gather of numpy buffers; gloo or nccl).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def minibatch_ranges(n_reads: int, minibatch: int) -> List[Tuple[int, int]]:
    return [(s, min(s + minibatch, n_reads)) for s in range(0, n_reads, minibatch)]


def shard_minibatches(n_reads: int, minibatch: int, rank: int, world: int) -> List[Tuple[int, int, int]]:
    """(minibatch index, first read, one-past-last read) of every minibatch owned by `rank`."""
    return [(i, a, b) for i, (a, b) in enumerate(minibatch_ranges(n_reads, minibatch)) if i % world == rank]


def local_read_indices(n_reads: int, minibatch: int, rank: int, world: int) -> np.ndarray:
    parts = [np.arange(a, b, dtype=np.int64) for _, a, b in shard_minibatches(n_reads, minibatch, rank, world)]
    return np.concatenate(parts) if parts else np.zeros(0, np.int64)


def gather_records(local: np.ndarray, n_reads: int, minibatch: int, rank: int, world: int, dist=None,
                   dst: int = 0) -> Optional[np.ndarray]:
    """Gather per-rank record arrays (structured dtype, local shard order) into global read order on `dst`."""
    if world == 1 or dist is None:
        return local
    import torch

    payload = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    sizes = [len(local_read_indices(n_reads, minibatch, r, world)) * local.dtype.itemsize for r in range(world)]
    bufs = [torch.empty(s, dtype=torch.uint8) for s in sizes] if rank == dst else None
    # gloo's gather needs equal sizes; pad to the largest shard
    big = max(sizes)
    padded = torch.zeros(big, dtype=torch.uint8)
    padded[: payload.numel()] = payload
    gathered = [torch.empty(big, dtype=torch.uint8) for _ in range(world)] if rank == dst else None
    dist.gather(padded, gathered, dst=dst)
    if rank != dst:
        return None
    out = np.zeros(n_reads, dtype=local.dtype)
    for r in range(world):
        idx = local_read_indices(n_reads, minibatch, r, world)
        out[idx] = gathered[r][: sizes[r]].numpy().view(local.dtype)
    del bufs
    return out


def detect_sharded(adc: np.ndarray, offsets: np.ndarray, full_lens: np.ndarray, calib_offset: np.ndarray,
                   calib_scale: np.ndarray, spc, model=None, minibatch_size: int = 1000, rank: int = 0,
                   world: int = 1, device: int = 0, dist=None,
                   detect_fn: Optional[Callable] = None):
    """Run this rank's minibatches on its GPU and gather the records on rank 0 (None elsewhere)."""
    n = int(np.asarray(full_lens).size)
    if detect_fn is None:
        from .detect import detect_reads

        def detect_fn(a, o, l, co, cs):  # noqa: E306
            return detect_reads(a, o, l, co, cs, spc, model=model, minibatch_size=minibatch_size, device=device,
                                return_records=True)[0]

    idx = local_read_indices(n, minibatch_size, rank, world)
    # compact this rank's reads into a contiguous ragged batch (minibatch boundaries are preserved because every
    # owned minibatch is complete except possibly the globally last one)
    lens = (np.asarray(offsets)[idx + 1] - np.asarray(offsets)[idx]).astype(np.int64)
    loc_off = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(lens, out=loc_off[1:])
    loc_adc = np.empty(int(loc_off[-1]), dtype=np.int16)
    for j, i in enumerate(idx):
        loc_adc[loc_off[j]: loc_off[j + 1]] = adc[offsets[i]: offsets[i + 1]]
    recs = detect_fn(loc_adc, loc_off, np.asarray(full_lens)[idx], np.asarray(calib_offset)[idx],
                     np.asarray(calib_scale)[idx])
    return gather_records(recs, n, minibatch_size, rank, world, dist)
