"""Device-side mirror of the reference's post-box plumbing (reference dloc/core/utils/utils.py:476-564 and the integer gating
of evaluation.py:82-103): same function names, arguments and return values, but the crop and the bicubic resize run in
liboetr_b200.so (oetr_crop_resize) on the images' GPU instead of D2H -> cv2.resize -> H2D.  No CPU or PyTorch fallback."""
import ctypes
import math

import torch

from oetr_b200 import cabi

MUL255, DIV255 = 1, 2


class _Job(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("channels", ctypes.c_int), ("src_h", ctypes.c_int),
                ("src_w", ctypes.c_int), ("x0", ctypes.c_int), ("y0", ctypes.c_int), ("x1", ctypes.c_int), ("y1", ctypes.c_int),
                ("new_w", ctypes.c_int), ("new_h", ctypes.c_int), ("flags", ctypes.c_int)]


def patch_resize(origin_w, origin_h, w, h, extractor_name):
    """reference dloc/core/utils/utils.py:476-494 (host arithmetic, unchanged semantics)."""
    if extractor_name != 'disk':
        if float(origin_w) / float(w) > float(origin_h) / float(h):
            ratio = float(origin_h) / float(h)
            new_w, new_h = ratio * float(w), origin_h
        else:
            ratio = float(origin_w) / float(w)
            new_w, new_h = origin_w, ratio * float(h)
        ratio = [[ratio, ratio]]
    else:
        ratio = [[float(origin_w) / float(w), float(origin_h) / float(h)]]
        new_w, new_h = origin_w, origin_h
    return ratio, int(new_w), int(new_h)


def crop_resize(jobs, device):
    """jobs: list of (image [C,H,W] fp32 CUDA tensor, (x0, y0, x1, y1), new_w, new_h, flags) -> list of [C,new_h,new_w]
    tensors, ONE kernel launch on torch's current stream."""
    if device.type != "cuda":
        raise cabi.OetrError(cabi.OETR_E_ARCH, "device %s: the crop / resize kernel has no CPU fallback" % (device,))
    lib = cabi.load_library()
    arr = (_Job * len(jobs))()
    outs, keep = [], []
    for i, (img, box, new_w, new_h, flags) in enumerate(jobs):
        img = img.contiguous().float()
        keep.append(img)
        c, h, w = img.shape
        out = torch.empty(c, new_h, new_w, dtype=torch.float32, device=device)
        outs.append(out)
        arr[i] = _Job(img.data_ptr(), out.data_ptr(), c, h, w, int(box[0]), int(box[1]), int(box[2]), int(box[3]), int(new_w),
                      int(new_h), int(flags))
    with torch.cuda.device(device):
        rc = lib.oetr_crop_resize(arr, len(jobs), ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    if rc != cabi.OETR_OK:
        raise cabi.OetrError(rc, (lib.oetr_crop_last_error() or b"").decode("utf-8", "replace"))
    return outs


def overlap_gate(bbox0, bbox1, dataset_name=''):
    """The integer gating of evaluation.py:86-103 on [1,4] box tensors (already multiplied by the scales): True when the
    overlap crops are used, False when the caller falls back to the full images."""
    b0, b1 = bbox0[0].int().tolist(), bbox1[0].int().tolist()               # one D2H of 8 integers
    bw0, bh0, bw1, bh1 = b0[2] - b0[0], b0[3] - b0[1], b1[2] - b1[0], b1[3] - b1[1]
    if min(bw0, bh0, bw1, bh1) <= 1:
        return False
    if dataset_name != 'pragueparks-val':
        return True
    return max(bw0 // bw1, bh0 // bh1, bw1 // bw0, bh1 // bh0) > 2.0


def tensor_overlap_crop(image1, bbox1, image2, bbox2, extractor_name, size_divisor=1):
    """reference dloc/core/utils/utils.py:510-564: image* [1,C,H,W] fp32 in [0,1] on a CUDA device, bbox* [1,4] ->
    (left [1,C,h1,w1], right [1,C,h2,w2], ratio1, ratio2).  Both crops go through one launch per resize pass."""
    b1, b2 = bbox1[0].int().tolist(), bbox2[0].int().tolist()
    origin_w1, origin_h1 = image1.shape[3], image1.shape[2]
    origin_w2, origin_h2 = image2.shape[3], image2.shape[2]
    w1, h1 = min(b1[2], origin_w1) - b1[0], min(b1[3], origin_h1) - b1[1]
    w2, h2 = min(b2[2], origin_w2) - b2[0], min(b2[3], origin_h2) - b2[1]
    if origin_w1 * origin_h1 >= origin_w2 * origin_h2:
        ow, oh = origin_w1, origin_h1
    else:
        ow, oh = origin_w2, origin_h2
    ratio1, new_w1, new_h1 = patch_resize(ow, oh, w1, h1, extractor_name)
    ratio2, new_w2, new_h2 = patch_resize(ow, oh, w2, h2, extractor_name)
    dev = image1.device
    two = size_divisor > 1
    left, right = crop_resize([(image1[0], b1, new_w1, new_h1, MUL255 | (0 if two else DIV255)),
                               (image2[0], b2, new_w2, new_h2, MUL255 | (0 if two else DIV255))], dev)
    if two:
        up = lambda v: math.ceil(v / size_divisor) * size_divisor
        left, right = crop_resize([(left, (0, 0, new_w1, new_h1), up(new_w1), up(new_h1), DIV255),
                                   (right, (0, 0, new_w2, new_h2), up(new_w2), up(new_h2), DIV255)], dev)
    return left[None], right[None], ratio1, ratio2
