"""Host-side driver of the CUDA neck (oetr_neck_*, include/oetr_b200.h): input_proj -> PatchMerging -> input_proj2 of the
reference's OETR.feature_extraction (src/model.py:116-124) for one image set.  PyTorch is plumbing (device memory,
streams); the arithmetic happens inside liboetr_b200.so.  No CPU or PyTorch fallback."""
import ctypes

import torch

from . import cabi
from .weights import NECK_PACKED_COUNT, pack_neck_weights


class NeckError(cabi.OetrError):
    pass


def _check(rc, lib):
    if rc != cabi.OETR_OK:
        raise NeckError(rc, (lib.oetr_neck_last_error() or b"").decode("utf-8", "replace"))


class NeckB200:
    """One handle = one immutable set of neck weights on one GPU."""

    def __init__(self, state_dict, device=None):
        self._lib = cabi.load_library()
        if not torch.cuda.is_available():
            raise NeckError(cabi.OETR_E_ARCH, "no CUDA device: the OETR neck has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise NeckError(cabi.OETR_E_ARCH, "device %s: the OETR neck has no CPU fallback" % (self.device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        packed = pack_neck_weights(state_dict)
        assert packed.size == NECK_PACKED_COUNT == self._lib.oetr_neck_packed_weight_count()
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(self._lib.oetr_neck_create(packed.ctypes.data_as(ctypes.c_void_p), packed.size,
                                              ctypes.byref(self._handle)), self._lib)
        self._ws = {}

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.oetr_neck_destroy(self._handle)
            self._handle = None

    __del__ = close

    @property
    def last_launch_count(self):
        return self._lib.oetr_neck_last_launch_count(self._handle)

    def _workspace(self, n, h, w):
        need = ctypes.c_size_t()
        _check(self._lib.oetr_neck_workspace_bytes(self._handle, n, h, w, ctypes.byref(need)), self._lib)
        key = torch.cuda.current_stream(self.device).cuda_stream          # one block per stream, as in OverlapHotPath
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need.value:
            ws = self._ws[key] = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return ws

    def forward(self, backbone_out, out=None):
        """backbone_out [n,1024,h,w] fp32 CUDA tensor (NCHW) -> feat [n,256,h//2,w//2] fp32.  Stream-ordered on torch's
        current stream; no synchronisation."""
        if backbone_out.dim() != 4 or backbone_out.shape[1] != 1024:
            raise ValueError("backbone features must be [n,1024,h,w], got %s" % (tuple(backbone_out.shape),))
        if backbone_out.device != self.device:
            raise ValueError("backbone features must live on %s" % self.device)
        x = backbone_out.contiguous().float()
        n, _, h, w = x.shape
        if out is None:
            out = torch.empty(n, 256, h // 2, w // 2, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._workspace(n, h, w)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _check(self._lib.oetr_neck_forward(self._handle, ctypes.c_void_p(x.data_ptr()), n, h, w,
                                               ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                               ctypes.c_void_p(stream)), self._lib)
        return out
