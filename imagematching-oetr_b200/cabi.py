"""ctypes binding of liboetr_b200.so (include/oetr_b200.h).  The CUDA library is mandatory: importing this
module never falls back to another implementation, and every entry point raises when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboetr_b200.so")

OETR_OK, OETR_E_ARG, OETR_E_SHAPE, OETR_E_ARCH, OETR_E_CUDA, OETR_E_NOMEM = 0, -1, -2, -3, -4, -5
ATTN_LINEAR, ATTN_FULL = 0, 1
PREC_FP32, PREC_FP16 = 0, 1

EXPORTS = (
    "oetr_abi_version", "oetr_last_error", "oetr_packed_weight_count", "oetr_create", "oetr_destroy",
    "oetr_workspace_bytes", "oetr_forward", "oetr_forward_masked", "oetr_last_launch_count", "oetr_poll_error", "oetr_forward_host",
    "oetr_forward_host_submit", "oetr_forward_host_wait",
    "oetr_profile_enable", "oetr_profile_read", "oetr_set_chunk_pairs",
    "oetr_selftest_tcgen05", "oetr_selftest_geometry", "oetr_debug_cycles",
    "oetr_gather_create", "oetr_gather_connect", "oetr_gather_submit", "oetr_gather_collect", "oetr_gather_destroy",
    "oetr_gather_last_error", "oetr_head_forward",
    "oetr_neck_packed_weight_count", "oetr_neck_create", "oetr_neck_destroy", "oetr_neck_workspace_bytes", "oetr_neck_forward",
    "oetr_neck_last_launch_count", "oetr_neck_geometry", "oetr_neck_last_error",
    "oetr_crop_resize", "oetr_crop_last_error",
    "oetr_sg_attention", "oetr_sg_transport_workspace_bytes", "oetr_sg_optimal_transport", "oetr_sg_last_error",
)
IPC_HANDLE_BYTES = 64


class OetrError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("liboetr_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen the in-tree library and declare the prototypes of include/oetr_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise OetrError(OETR_E_ARCH, "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "(there is no CPU or PyTorch fallback for the hot path)" % path)
    lib = ctypes.CDLL(path)
    c = ctypes
    f32p, vp = c.POINTER(c.c_float), c.c_void_p
    lib.oetr_abi_version.restype = c.c_int
    lib.oetr_last_error.restype = c.c_char_p
    lib.oetr_packed_weight_count.restype = c.c_size_t
    lib.oetr_create.restype = c.c_int
    lib.oetr_create.argtypes = [vp, c.c_size_t, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(vp)]
    lib.oetr_destroy.restype = c.c_int
    lib.oetr_destroy.argtypes = [vp]
    lib.oetr_workspace_bytes.restype = c.c_int
    lib.oetr_workspace_bytes.argtypes = [vp, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_size_t)]
    lib.oetr_forward.restype = c.c_int
    lib.oetr_forward.argtypes = [vp, vp, vp] + [c.c_int] * 10 + [vp] * 6 + [vp, c.c_size_t, vp]
    lib.oetr_forward_masked.restype = c.c_int
    lib.oetr_forward_masked.argtypes = [vp, vp, vp, vp, vp] + [c.c_int] * 10 + [vp] * 6 + [vp, c.c_size_t, vp]
    lib.oetr_last_launch_count.restype = c.c_int
    lib.oetr_last_launch_count.argtypes = [vp]
    lib.oetr_set_chunk_pairs.restype = c.c_int
    lib.oetr_set_chunk_pairs.argtypes = [vp, c.c_int]
    lib.oetr_profile_enable.restype = c.c_int
    lib.oetr_profile_enable.argtypes = [vp, c.c_int]
    lib.oetr_profile_read.restype = c.c_int
    lib.oetr_profile_read.argtypes = [vp, f32p, c.POINTER(c.c_int)]
    lib.oetr_poll_error.restype = c.c_int
    lib.oetr_poll_error.argtypes = [vp]
    lib.oetr_forward_host.restype = c.c_int
    lib.oetr_forward_host.argtypes = [vp, vp, vp] + [c.c_int] * 10 + [vp, vp, vp]
    lib.oetr_forward_host_submit.restype = c.c_int
    lib.oetr_forward_host_submit.argtypes = [vp, vp, vp] + [c.c_int] * 10 + [vp, c.POINTER(c.c_int)]
    lib.oetr_forward_host_wait.restype = c.c_int
    lib.oetr_forward_host_wait.argtypes = [vp, c.c_int, vp, vp]
    lib.oetr_selftest_geometry.restype = c.c_int
    lib.oetr_selftest_geometry.argtypes = [c.c_int] * 5 + [c.POINTER(c.c_int)]
    lib.oetr_head_forward.restype = c.c_int
    lib.oetr_head_forward.argtypes = [vp] * 7 + [c.c_int] * 10 + [vp] * 4 + [vp, c.c_size_t, vp]
    lib.oetr_gather_create.restype = c.c_int
    lib.oetr_gather_create.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(vp), vp]
    lib.oetr_gather_connect.restype = c.c_int
    lib.oetr_gather_connect.argtypes = [vp, vp]
    lib.oetr_gather_submit.restype = c.c_int
    lib.oetr_gather_submit.argtypes = [vp, vp, vp, vp]
    lib.oetr_gather_collect.restype = c.c_int
    lib.oetr_gather_collect.argtypes = [vp, vp, vp]
    lib.oetr_gather_destroy.restype = c.c_int
    lib.oetr_gather_destroy.argtypes = [vp]
    lib.oetr_gather_last_error.restype = c.c_char_p
    lib.oetr_neck_packed_weight_count.restype = c.c_size_t
    lib.oetr_neck_create.restype = c.c_int
    lib.oetr_neck_create.argtypes = [vp, c.c_size_t, c.POINTER(vp)]
    lib.oetr_neck_destroy.restype = c.c_int
    lib.oetr_neck_destroy.argtypes = [vp]
    lib.oetr_neck_workspace_bytes.restype = c.c_int
    lib.oetr_neck_workspace_bytes.argtypes = [vp, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_size_t)]
    lib.oetr_neck_forward.restype = c.c_int
    lib.oetr_neck_forward.argtypes = [vp, vp, c.c_int, c.c_int, c.c_int, vp, vp, c.c_size_t, vp]
    lib.oetr_neck_last_launch_count.restype = c.c_int
    lib.oetr_neck_last_launch_count.argtypes = [vp]
    lib.oetr_neck_geometry.restype = c.c_int
    lib.oetr_neck_geometry.argtypes = [c.c_int] * 4 + [c.POINTER(c.c_int)]
    lib.oetr_neck_last_error.restype = c.c_char_p
    lib.oetr_crop_resize.restype = c.c_int
    lib.oetr_crop_resize.argtypes = [vp, c.c_int, vp]
    lib.oetr_crop_last_error.restype = c.c_char_p
    lib.oetr_sg_attention.restype = c.c_int
    lib.oetr_sg_attention.argtypes = [vp, vp, vp, vp, c.c_int, c.c_int, c.c_int, c.c_int, vp]
    lib.oetr_sg_transport_workspace_bytes.restype = c.c_size_t
    lib.oetr_sg_transport_workspace_bytes.argtypes = [c.c_int, c.c_int, c.c_int]
    lib.oetr_sg_optimal_transport.restype = c.c_int
    lib.oetr_sg_optimal_transport.argtypes = [vp, c.c_float, c.c_int, vp, c.c_int, c.c_int, c.c_int, vp, c.c_size_t, vp]
    lib.oetr_sg_last_error.restype = c.c_char_p
    lib.oetr_debug_cycles.restype = c.c_int
    lib.oetr_debug_cycles.argtypes = [c.POINTER(c.c_ulonglong), c.c_int, c.c_int]
    lib.oetr_selftest_tcgen05.restype = c.c_int
    lib.oetr_selftest_tcgen05.argtypes = [f32p, c.c_int]
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(rc, lib=None):
    if rc != OETR_OK:
        lib = lib or load_library()
        raise OetrError(rc, (lib.oetr_last_error() or b"").decode("utf-8", "replace"))
