"""world_size-2 gloo test of the batch sharding + box all-gather used for N>1 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oetr_b200.distributed import BoxGather, ShardedOverlapEstimator, shard_range


def test_shard_ranges_cover_batch():
    for batch in (0, 1, 5, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_compute(f1, f2):
    # a per-pair function of the inputs, so ordering mistakes are visible
    s1 = f1.sum(dim=(1, 2, 3))
    s2 = f2.sum(dim=(1, 2, 3))
    return torch.stack([s1, s1 + 1, s1 + 2, s1 + 3], 1), torch.stack([s2, s2 - 1, s2 - 2, s2 - 3], 1)


def _worker(rank, world, port, batch, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        f1 = torch.randn(batch, 4, 3, 3, generator=g)
        f2 = torch.randn(batch, 4, 2, 3, generator=g)
        b1, b2 = ShardedOverlapEstimator(_fake_compute)(f1, f2)
        e1, e2 = _fake_compute(f1, f2)
        ok[rank] = int(torch.equal(b1, e1) and torch.equal(b2, e2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8])
def test_gather_two_ranks(batch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, batch, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


def _gather_worker(rank, world, port, pairs, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = BoxGather(pairs, "cpu", slots=3)
        assert g.mode == "collective"
        good = True
        for step in range(7):                               # more steps than slots: the ring wraps
            b1 = torch.arange(pairs * 4, dtype=torch.float32).view(pairs, 4) + 1000 * rank + 10000 * step
            g.submit(b1, -b1)
            got = g.result()
            for r in range(world):
                want = torch.arange(pairs * 4, dtype=torch.float32).view(pairs, 4) + 1000 * r + 10000 * step
                good &= bool(torch.equal(got[r * pairs:(r + 1) * pairs, 0], want))
                good &= bool(torch.equal(got[r * pairs:(r + 1) * pairs, 1], -want))
        ok[rank] = int(good)
    finally:
        dist.destroy_process_group()


def test_box_gather_ring_two_ranks_gloo():
    """BoxGather (collective mode on CPU): every rank ends every step with the boxes of all ranks in global order."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_gather_worker, args=(2, port, 3, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
