"""Host-side logic of the multi-GPU path on CPU: world size 2 over gloo, a stub detector in place of the GPU call."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adapted_b200.pipeline import detect_sharded, local_read_indices, shard_minibatches

REC = np.dtype([("read", "<i8"), ("first_adc", "<i4"), ("n", "<i4")])


def test_shards_partition_the_minibatches():
    n, mb = 10431, 1000
    seen = np.zeros(n, int)
    for world in (1, 2, 4, 8):
        seen[:] = 0
        for r in range(world):
            for i, a, b in shard_minibatches(n, mb, r, world):
                assert i % world == r and b - a <= mb
                seen[a:b] += 1
        assert (seen == 1).all()
    assert local_read_indices(2500, 1000, 1, 2).tolist() == list(range(1000, 2000))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, mb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    lens = rng.integers(5, 40, size=n)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    adc = rng.integers(-500, 500, size=int(offsets[-1])).astype(np.int16)
    full = lens.astype(np.int32)
    tag = np.arange(n, dtype=np.float32)  # rides in calib_offset so the stub can recover the global read index

    def stub(a, o, l, co, cs):
        out = np.zeros(l.size, REC)
        out["read"] = co.astype(np.int64)
        out["first_adc"] = a[o[:-1]]
        out["n"] = (o[1:] - o[:-1]).astype(np.int32)
        return out

    res = detect_sharded(adc, offsets, full, tag, np.ones(n, np.float32), None, minibatch_size=mb, rank=rank,
                         world=world, dist=dist, detect_fn=stub)
    if rank == 0:
        ok = (res["read"] == np.arange(n)).all() and (res["first_adc"] == adc[offsets[:-1]]).all() and (res["n"] == lens).all()
        q.put(bool(ok))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_gather_restores_read_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 2350, 100, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
