"""2-rank NCCL test (needs 2 GPUs; skipped otherwise): the boxes the CUDA hot path produces on each rank arrive on every
rank in global pair order, through BOTH gather transports (peer stores over CUDA IPC, and the NCCL collective), and equal
the CPU oracle's boxes for the same global batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ok):
    import torch.distributed as dist

    import oetr_b200
    from oetr_b200 import weights
    from oetr_b200.distributed import BoxGather, shard_range
    from oracle import oetr_oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        W = weights.synthetic_hot_path_weights(0)
        batch, per = 6, 3
        f1 = weights.synthetic_features(batch, 8, 10, seed=41, tag="g1")
        f2 = weights.synthetic_features(batch, 6, 7, seed=41, tag="g2")
        hw1, hw2 = (256, 320), (192, 224)
        want = orc.hot_path(W, f1, f2, hw1, hw2, clamp=False)
        want = np.stack([want["box1_raw"], want["box2_raw"]], axis=1)                # [batch, 2, 4]
        s, e = shard_range(batch, rank, world)
        hot = oetr_b200.OverlapHotPath(W, precision="fp16", device=dev)
        good = True
        for mode in ("peer", "collective"):
            g = BoxGather(per, dev, mode=mode, slots=3)
            good &= g.mode == mode
            for step in range(5):                                                   # the slot ring wraps
                b1, b2 = hot.forward(torch.from_numpy(f1[s:e]).to(dev), torch.from_numpy(f2[s:e]).to(dev), hw1, hw2, clamp=False)
                g.submit(b1 + step, b2 + step)
                g.wait()
                got = g.result().cpu().numpy() - step
                good &= bool(np.abs(got - want).max() / 320 < 1e-3)
                good &= bool(np.array_equal(got[s:e, 0], b1.cpu().numpy()))          # own slice: bit-exact
            g.close()
        hot.close()
        ok[rank] = int(good)
    finally:
        dist.destroy_process_group()


def test_gathered_boxes_two_ranks_nccl_and_peer_stores():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
