"""Multi-GPU plumbing: pairs are independent in eval mode, so the batch is split contiguously over ranks
(one process per GPU, replicated weights, no data-path collective) and only the [B,2,4] fp32 boxes are
all-gathered at the end.  The reference has no inference-time communication."""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world_size):
    """Contiguous [start, stop) of `batch` pairs owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(batch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_boxes(box1, box2, batch, group=None):
    """Every rank contributes its shard's boxes ([n_r,4] each) and receives the full ([batch,4], [batch,4]) in
    global pair order.  Uneven shards are padded to the largest shard for the fixed-size collective."""
    world = dist.get_world_size(group)
    if world == 1:
        return box1, box2
    cap = -(-batch // world)
    local = torch.zeros(cap, 2, 4, dtype=torch.float32, device=box1.device)
    n = box1.shape[0]
    local[:n, 0] = box1
    local[:n, 1] = box2
    gathered = torch.empty(world * cap, 2, 4, dtype=torch.float32, device=box1.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    gathered = gathered.view(world, cap, 2, 4)
    parts = []
    for r in range(world):
        s, e = shard_range(batch, r, world)
        parts.append(gathered[r, : e - s])
    out = torch.cat(parts, dim=0)
    return out[:, 0].contiguous(), out[:, 1].contiguous()


class ShardedOverlapEstimator:
    """Runs `compute(feat1_shard, feat2_shard) -> (box1, box2)` on this rank's slice of a global batch and
    returns the gathered boxes of the whole batch on every rank."""

    def __init__(self, compute, group=None):
        self.compute = compute
        self.group = group

    def __call__(self, feat1, feat2):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        batch = feat1.shape[0]
        s, e = shard_range(batch, rank, world)
        box1, box2 = self.compute(feat1[s:e], feat2[s:e])
        if world == 1:
            return box1, box2
        return all_gather_boxes(box1, box2, batch, self.group)
