"""Host-side driver of the CUDA hot path: owns an oetr_handle, a torch-allocated workspace, and exposes
`forward(feat1, feat2, ...)` replacing the reference's feature_correlation + center_estimation +
size_regression + box_tlbr_to_xyxy (reference src/model.py:240-250).  PyTorch is plumbing here (device memory
and streams); all arithmetic happens inside liboetr_b200.so."""
import ctypes

import numpy as np
import torch

from . import cabi
from .weights import PACKED_COUNT, pack_hot_path_weights

_ATTN = {"linear": cabi.ATTN_LINEAR, "full": cabi.ATTN_FULL}
_PREC = {"fp32": cabi.PREC_FP32, "fp16": cabi.PREC_FP16}


class OverlapHotPath:
    """One handle = one immutable set of hot-path weights on one GPU."""

    def __init__(self, state_dict, attention="linear", precision="fp16", max_shape=(100, 100), device=None):
        if attention not in _ATTN:
            raise ValueError("attention %r not in %s" % (attention, sorted(_ATTN)))
        if precision not in _PREC:
            raise ValueError("precision %r not in %s" % (precision, sorted(_PREC)))
        self._lib = cabi.load_library()
        if not torch.cuda.is_available():
            raise cabi.OetrError(cabi.OETR_E_ARCH, "no CUDA device: the OETR hot path has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise cabi.OetrError(cabi.OETR_E_ARCH, "device %s: the OETR hot path has no CPU fallback" % (self.device,))
        if self.device.index is None:                      # 'cuda' -> the current device, so tensor.device compares equal
            self.device = torch.device("cuda", torch.cuda.current_device())
        packed = pack_hot_path_weights(state_dict)
        assert packed.size == PACKED_COUNT == self._lib.oetr_packed_weight_count()
        self.attention, self.precision, self.max_shape = attention, precision, tuple(max_shape)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            cabi.check(self._lib.oetr_create(packed.ctypes.data_as(ctypes.c_void_p), packed.size, 0,
                                             _ATTN[attention], _PREC[precision], int(max_shape[0]),
                                             int(max_shape[1]), ctypes.byref(self._handle)), self._lib)
        self._ws = {}                                      # one workspace per CUDA stream (concurrent forwards)
        self._inflight = {}

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.oetr_destroy(self._handle)
            self._handle = None

    __del__ = close

    def _workspace(self, batch, hf1, wf1, hf2, wf2):
        need = ctypes.c_size_t()
        cabi.check(self._lib.oetr_workspace_bytes(self._handle, batch, hf1, wf1, hf2, wf2, ctypes.byref(need)),
                   self._lib)
        # one block per stream: forwards on two streams of one handle must not share activations, and a block is
        # allocated (and later freed) under the stream that uses it, so the caching allocator's reuse is stream-safe
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need.value:
            ws = self._ws[key] = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return ws

    def set_chunk_pairs(self, pairs):
        """Pairs per concurrently scheduled sub-batch of the fp16 path (0 = never split; < 0 = automatic, the default:
        about 56 encoder tiles per sub-batch, i.e. 8 pairs at 640x640)."""
        cabi.check(self._lib.oetr_set_chunk_pairs(self._handle, int(pairs)), self._lib)
        self._ws = {}

    def profile(self, enable=True):
        cabi.check(self._lib.oetr_profile_enable(self._handle, int(bool(enable))), self._lib)

    def profile_read(self):
        """(average k_enc launch duration in ms, launches) since the last read; synchronises."""
        ms, n = ctypes.c_float(), ctypes.c_int()
        cabi.check(self._lib.oetr_profile_read(self._handle, ctypes.byref(ms), ctypes.byref(n)), self._lib)
        return ms.value, n.value

    def poll_error(self):
        """Synchronise and raise if an earlier forward failed asynchronously (tests / debugging)."""
        cabi.check(self._lib.oetr_poll_error(self._handle), self._lib)

    @property
    def last_launch_count(self):
        return self._lib.oetr_last_launch_count(self._handle)

    def forward(self, feat1, feat2, img_hw1, img_hw2, clamp=True, debug=False, mask1=None, mask2=None):
        """feat1 [B,256,hf1,wf1], feat2 [B,256,hf2,wf2]: fp32 CUDA tensors (NCHW).  Returns (box1, box2) [B,4]
        xyxy pixels; with debug=True also a dict of the stage boundaries (hs, memory, cxy, tlbr).
        mask1 [B,hf1,wf1], mask2 [B,hf2,wf2]: the optional masks of the reference's forward_dummy (both or neither).
        Stream-ordered on torch's current stream; no synchronisation."""
        if feat1.dim() != 4 or feat2.dim() != 4 or feat1.shape[1] != 256 or feat2.shape[1] != 256:
            raise ValueError("features must be [B,256,h,w], got %s and %s" % (tuple(feat1.shape), tuple(feat2.shape)))
        if feat1.shape[0] != feat2.shape[0]:
            raise ValueError("batch mismatch: %d vs %d" % (feat1.shape[0], feat2.shape[0]))
        if feat1.device != self.device or feat2.device != self.device:
            raise ValueError("features must live on %s" % self.device)
        feat1 = feat1.contiguous().float()
        feat2 = feat2.contiguous().float()
        b, _, hf1, wf1 = feat1.shape
        _, _, hf2, wf2 = feat2.shape
        if (mask1 is None) != (mask2 is None):
            raise ValueError("give both masks or neither")
        if mask1 is not None:
            if tuple(mask1.shape) != (b, hf1, wf1) or tuple(mask2.shape) != (b, hf2, wf2):
                raise ValueError("masks must be [B,hf,wf] like the feature maps, got %s and %s" % (
                    tuple(mask1.shape), tuple(mask2.shape)))
            mask1 = mask1.to(device=self.device, dtype=torch.float32).contiguous()
            mask2 = mask2.to(device=self.device, dtype=torch.float32).contiguous()
        box1 = torch.empty(b, 4, dtype=torch.float32, device=self.device)
        box2 = torch.empty(b, 4, dtype=torch.float32, device=self.device)
        dbg = None
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else ctypes.c_void_p(0)
        if debug:
            dbg = dict(hs=torch.empty(2, b, 256, device=self.device),
                       memory=torch.empty(b * (hf1 * wf1 + hf2 * wf2), 256, device=self.device),
                       cxy=torch.empty(2, b, 2, device=self.device), tlbr=torch.empty(2, b, 4, device=self.device))
        with torch.cuda.device(self.device):
            ws = self._workspace(b, hf1, wf1, hf2, wf2)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self._lib.oetr_forward_masked(
                self._handle, ptr(feat1), ptr(feat2), ptr(mask1), ptr(mask2), b, hf1, wf1, hf2, wf2,
                int(img_hw1[0]), int(img_hw1[1]), int(img_hw2[0]), int(img_hw2[1]), int(bool(clamp)),
                ptr(box1), ptr(box2),
                ptr(dbg["hs"]) if debug else None, ptr(dbg["memory"]) if debug else None,
                ptr(dbg["cxy"]) if debug else None, ptr(dbg["tlbr"]) if debug else None,
                ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(stream)), self._lib)
        if not debug:
            return box1, box2
        l1 = hf1 * wf1
        out = dict(hs1=dbg["hs"][0], hs2=dbg["hs"][1], cxy1=dbg["cxy"][0], cxy2=dbg["cxy"][1],
                   tlbr1=dbg["tlbr"][0], tlbr2=dbg["tlbr"][1],
                   memory1=dbg["memory"][: b * l1].view(b, l1, 256),
                   memory2=dbg["memory"][b * l1:].view(b, hf2 * wf2, 256))
        return box1, box2, out

    def head(self, hs1, hs2, memory1, memory2, hf1, wf1, hf2, wf2, img_hw1, img_hw2, clamp=True, mask1=None, mask2=None):
        """The overlap head alone (oetr_head_forward): hs [B,1,256] or [B,256], memory [B,L,256] CUDA fp32 tensors ->
        (box1, box2 [B,4], cxy1, cxy2 [B,2], tlbr1, tlbr2 [B,4]).  Serves the reference's stage-wise signatures
        (center_estimation / size_regression, src/model.py:145-191)."""
        b = memory1.shape[0]
        f = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        hs1, hs2, memory1, memory2 = f(hs1).reshape(b, 256), f(hs2).reshape(b, 256), f(memory1), f(memory2)
        if tuple(memory1.shape) != (b, hf1 * wf1, 256) or tuple(memory2.shape) != (b, hf2 * wf2, 256):
            raise ValueError("memory must be [B,hf*wf,256], got %s and %s" % (tuple(memory1.shape), tuple(memory2.shape)))
        if (mask1 is None) != (mask2 is None):
            raise ValueError("give both masks or neither")
        if mask1 is not None:
            mask1, mask2 = f(mask1.float()), f(mask2.float())
        box1 = torch.empty(b, 4, dtype=torch.float32, device=self.device)
        box2 = torch.empty(b, 4, dtype=torch.float32, device=self.device)
        cxy = torch.empty(2, b, 2, dtype=torch.float32, device=self.device)
        tlbr = torch.empty(2, b, 4, dtype=torch.float32, device=self.device)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() else ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            ws = self._workspace(b, hf1, wf1, hf2, wf2)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self._lib.oetr_head_forward(
                self._handle, ptr(memory1), ptr(memory2), ptr(hs1), ptr(hs2), ptr(mask1), ptr(mask2), b, hf1, wf1, hf2, wf2,
                int(img_hw1[0]), int(img_hw1[1]), int(img_hw2[0]), int(img_hw2[1]), int(bool(clamp)), ptr(box1), ptr(box2),
                ptr(cxy), ptr(tlbr), ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(stream)), self._lib)
        return box1, box2, cxy[0], cxy[1], tlbr[0], tlbr[1]

    def submit_host(self, feat1, feat2, img_hw1, img_hw2, clamp=True):
        """Queue a host-buffer request (oetr_forward_host_submit) and return a ticket for `wait_host`.  At most four
        requests may be in flight; the copies of one overlap the compute of the others."""
        a1 = np.ascontiguousarray(feat1.numpy() if hasattr(feat1, "numpy") else feat1, dtype=np.float32)
        a2 = np.ascontiguousarray(feat2.numpy() if hasattr(feat2, "numpy") else feat2, dtype=np.float32)
        if a1.ndim != 4 or a2.ndim != 4 or a1.shape[1] != 256 or a2.shape[1] != 256 or a1.shape[0] != a2.shape[0]:
            raise ValueError("features must be [B,256,h,w] with equal B, got %s and %s" % (a1.shape, a2.shape))
        b, _, hf1, wf1 = a1.shape
        _, _, hf2, wf2 = a2.shape
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        ticket = ctypes.c_int(-1)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self._lib.oetr_forward_host_submit(
                self._handle, vp(a1), vp(a2), b, hf1, wf1, hf2, wf2, int(img_hw1[0]), int(img_hw1[1]),
                int(img_hw2[0]), int(img_hw2[1]), int(bool(clamp)), ctypes.c_void_p(stream), ctypes.byref(ticket)),
                self._lib)
        self._inflight[ticket.value] = (a1, a2, b)            # keeps the feature buffers alive until the wait
        return ticket.value

    def wait_host(self, ticket):
        """Block until the request has finished; returns (box1, box2) numpy [B,4]."""
        a1, a2, b = self._inflight.pop(ticket)
        box1 = np.empty((b, 4), dtype=np.float32)
        box2 = np.empty((b, 4), dtype=np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        with torch.cuda.device(self.device):
            cabi.check(self._lib.oetr_forward_host_wait(self._handle, ticket, vp(box1), vp(box2)), self._lib)
        return box1, box2

    def forward_host(self, feat1, feat2, img_hw1, img_hw2, clamp=True):
        """Host-buffer entry (numpy fp32 arrays or CPU tensors in, numpy boxes out): H2D copies, hot path, D2H
        copies and the wait for completion all inside the library (oetr_forward_host = submit + wait)."""
        return self.wait_host(self.submit_host(feat1, feat2, img_hw1, img_hw2, clamp))
