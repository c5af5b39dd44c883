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


class BoxGather:
    """The [pairs,4]+[pairs,4] boxes of every rank on every rank, once per step, OFF the compute streams.

    mode "peer" (default on CUDA with world > 1): liboetr_b200's oetr_gather_* -- every rank stores its boxes straight
    into its peers' buffers over NVLink (CUDA IPC) and a flag-wait kernel collects them; no collective, no rendezvous,
    no NCCL kernel competing for an SM.  mode "collective": `all_gather_into_tensor` (NCCL / gloo) -- the portable
    fallback.  Either way the work is issued on a dedicated communication stream that waits for the producing stream by
    event, so the lane that produced the boxes goes straight on to its next forward; `join(stream)` / `wait()` are the
    only synchronisation points.  Results land in a ring of output tensors (`result()` = the most recent one)."""

    def __init__(self, pairs, device, group=None, mode="auto", slots=8):
        self.pairs, self.device, self.group = int(pairs), torch.device(device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.slots = slots
        self.cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.out = [torch.zeros(self.world * self.pairs, 2, 4, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self.n = 0
        self._handle = None
        self._stage = [torch.empty(self.pairs, 2, 4, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self.mode = "collective"
        self.fallback_reason = None
        if mode in ("auto", "peer") and self.cuda and self.world > 1:
            try:
                self._connect_peers()
                self.mode = "peer"
            except Exception as e:                      # no P2P / IPC on this box: the collective still works
                if mode == "peer":
                    raise
                self.fallback_reason = repr(e)

    def _connect_peers(self):
        import ctypes

        from . import cabi
        lib = cabi.load_library()
        h = ctypes.c_void_p()
        mine = ctypes.create_string_buffer(cabi.IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            rc = lib.oetr_gather_create(self.world, self.rank, self.pairs, self.slots, ctypes.byref(h), mine)
        if rc != 0:
            raise RuntimeError("oetr_gather_create: " + (lib.oetr_gather_last_error() or b"").decode())
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=self.group)
        blob = b"".join(handles)
        with torch.cuda.device(self.device):
            rc = lib.oetr_gather_connect(h, blob)
        ok = torch.tensor([1 if rc == 0 else 0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)        # all ranks agree on the mode
        if int(ok.item()) == 0:
            err = (lib.oetr_gather_last_error() or b"").decode()
            lib.oetr_gather_destroy(h)
            raise RuntimeError("oetr_gather_connect failed on some rank: " + err)
        self._lib, self._handle = lib, h

    def submit(self, box1, box2):
        """Queue the gather of this step's boxes (tensors on `device`, produced on the CURRENT stream)."""
        slot = self.n % self.slots
        self.n += 1
        if self.world == 1:
            self.out[slot][:, 0], self.out[slot][:, 1] = box1, box2
            return
        if not self.cuda:
            local = torch.stack([box1, box2], dim=1).contiguous()
            dist.all_gather_into_tensor(self.out[slot], local, group=self.group)
            return
        import ctypes
        cur = torch.cuda.current_stream(self.device)
        self.comm.wait_stream(cur)
        box1.record_stream(self.comm)
        box2.record_stream(self.comm)
        with torch.cuda.stream(self.comm):
            if self.mode == "peer":
                s = ctypes.c_void_p(self.comm.cuda_stream)
                b1, b2 = box1.contiguous(), box2.contiguous()
                rc = self._lib.oetr_gather_submit(self._handle, ctypes.c_void_p(b1.data_ptr()), ctypes.c_void_p(b2.data_ptr()), s)
                if rc == 0:
                    rc = self._lib.oetr_gather_collect(self._handle, ctypes.c_void_p(self.out[slot].data_ptr()), s)
                if rc != 0:
                    raise RuntimeError("oetr_gather: " + (self._lib.oetr_gather_last_error() or b"").decode())
            else:
                st = self._stage[slot]
                st[:, 0], st[:, 1] = box1, box2
                dist.all_gather_into_tensor(self.out[slot], st, group=self.group)

    def submit_host(self, box1, box2, pinned):
        """Host boxes (numpy [pairs,4] each; `pinned` = a pinned [pairs,2,4] staging tensor) -> device -> gather."""
        if self.world == 1:
            return
        pinned[:, 0] = torch.from_numpy(box1)
        pinned[:, 1] = torch.from_numpy(box2)
        with torch.cuda.stream(self.comm):
            dev = pinned.to(self.device, non_blocking=True)
        self.comm.synchronize()                          # `pinned` is reused by the caller's next step
        self.submit(dev[:, 0], dev[:, 1])

    def join(self, stream):
        if self.comm is not None:
            stream.wait_stream(self.comm)

    def wait(self):
        if self.comm is not None:
            self.comm.synchronize()

    def result(self):
        return self.out[(self.n - 1) % self.slots]

    def close(self):
        if self._handle is not None:
            self.wait()
            if dist.is_initialized():
                dist.barrier(group=self.group)           # no peer is still writing into this rank's buffer
            self._lib.oetr_gather_destroy(self._handle)
            self._handle = None
