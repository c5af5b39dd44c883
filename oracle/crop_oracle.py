"""CPU oracle of the post-box plumbing -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy,
  * the integer gating of the overlap boxes          evaluation.py:82-103
  * patch_resize / tensor_overlap_crop               dloc/core/utils/utils.py:476-564
  * cv2.resize(float32, INTER_CUBIC), the bicubic the reference calls (third-party: opencv-python, not vendored in the
    reference tree; 4.13.0 in this image).  Published algorithm (imgproc/resize.cpp, `ResizeFunc` for CV_32F / cubic):
    per destination index d: f = (d + 0.5) * (src / dst) - 0.5, s = floor(f), t = f - s, taps s-1 .. s+2 with source
    indices CLAMPED to the image, Keys weights with A = -0.75
        w0 = ((A (t+1) - 5A)(t+1) + 8A)(t+1) - 4A,  w1 = ((A+2) t - (A+3)) t t + 1,
        w2 = ((A+2)(1-t) - (A+3))(1-t)(1-t) + 1,    w3 = 1 - w0 - w1 - w2,
    horizontal pass first, then vertical, pixel arithmetic in float32; equal sizes = plain copy.
Parity status: pinned against cv2 itself where it is importable (tests/test_crop_oracle.py) and against committed
fixtures made with it (tests/golden/crop_*.npz).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
"""
import math

import numpy as np

A64 = -0.75


def cubic_taps(src, dst):
    """(index [dst,4] int64 clamped, weight [dst,4] float32) of cv2's INTER_CUBIC along one axis.  The optimised build of
    cv2 4.13 (the default code path; cv2.setUseOptimized(False) differs by up to 5e-3 on a 0..255 image) evaluates the
    source position, its fraction and the four weights in double and rounds the weights to float32 -- established by
    probing cv2 in this container (tests/test_crop_oracle.py keeps the comparison)."""
    scale = 1.0 / (float(dst) / float(src))                    # resize.cpp: scale_x = 1. / inv_scale_x
    f = (np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5
    s = np.floor(f).astype(np.int64)
    t = f - s
    w = np.empty((dst, 4), dtype=np.float64)
    t1 = t + 1.0
    w[:, 0] = ((A64 * t1 - 5.0 * A64) * t1 + 8.0 * A64) * t1 - 4.0 * A64
    w[:, 1] = ((A64 + 2.0) * t - (A64 + 3.0)) * t * t + 1.0
    u = 1.0 - t
    w[:, 2] = ((A64 + 2.0) * u - (A64 + 3.0)) * u * u + 1.0
    w[:, 3] = 1.0 - w[:, 0] - w[:, 1] - w[:, 2]
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, src - 1)
    return idx, w.astype(np.float32)


def resize_cubic(img, new_w, new_h):
    """cv2.resize(img.astype('float32'), (new_w, new_h), interpolation=cv2.INTER_CUBIC); img [H,W] or [H,W,C]."""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    if (new_w, new_h) == (w, h):
        return img.copy()
    ix, wx = cubic_taps(w, new_w)
    iy, wy = cubic_taps(h, new_h)
    ex = (slice(None),) * 2 + (None,) * (img.ndim - 2)
    rows = np.zeros((h, new_w) + img.shape[2:], dtype=np.float32)
    for k in range(4):
        rows += img[:, ix[:, k]] * wx[None, :, k][ex]
    out = np.zeros((new_h, new_w) + img.shape[2:], dtype=np.float32)
    for k in range(4):
        out += rows[iy[:, k]] * wy[:, k][:, None][ex]
    return out


def patch_resize(origin_w, origin_h, w, h, extractor_name):
    """dloc/core/utils/utils.py:476-494"""
    if extractor_name != "disk":
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


def int_box(bbox):
    """torch's .int() on a float box: truncation toward zero."""
    return [int(np.trunc(v)) for v in np.asarray(bbox, dtype=np.float32).reshape(-1)[:4]]


def overlap_gate(bbox0, bbox1, dataset_name=""):
    """evaluation.py:86-103: True when the crops are used.  bbox* [4] float (already multiplied by the scales)."""
    b0, b1 = int_box(bbox0), int_box(bbox1)
    bw0, bh0, bw1, bh1 = b0[2] - b0[0], b0[3] - b0[1], b1[2] - b1[0], b1[3] - b1[1]
    if min(bw0, bh0, bw1, bh1) <= 1:
        return False
    if dataset_name != "pragueparks-val":
        return True
    fd = lambda a, b: math.floor(a / b)
    return max(fd(bw0, bw1), fd(bh0, bh1), fd(bw1, bw0), fd(bh1, bh0)) > 2.0


def tensor_overlap_crop(image1, bbox1, image2, bbox2, extractor_name, size_divisor=1):
    """dloc/core/utils/utils.py:510-564 on numpy arrays: image* [1,C,H,W] float32 in [0,1], bbox* [1,4] ->
    (left [1,C,h1,w1], right [1,C,h2,w2], ratio1, ratio2)."""
    b1, b2 = int_box(bbox1), int_box(bbox2)
    origin_w1, origin_h1 = image1.shape[3], image1.shape[2]
    origin_w2, origin_h2 = image2.shape[3], image2.shape[2]
    left = image1[0, :, b1[1]:b1[3], b1[0]:b1[2]]
    right = image2[0, :, b2[1]:b2[3], b2[0]:b2[2]]
    w1, h1 = left.shape[2], left.shape[1]
    w2, h2 = right.shape[2], right.shape[1]
    if origin_w1 * origin_h1 >= origin_w2 * origin_h2:
        ow, oh = origin_w1, origin_h1
    else:
        ow, oh = origin_w2, origin_h2
    ratio1, nw1, nh1 = patch_resize(ow, oh, w1, h1, extractor_name)
    ratio2, nw2, nh2 = patch_resize(ow, oh, w2, h2, extractor_name)
    outs = []
    for crop, nw, nh in ((left, nw1, nh1), (right, nw2, nh2)):
        cv = np.transpose(crop, (1, 2, 0)).astype(np.float32) * np.float32(255)
        cv = resize_cubic(cv, nw, nh)
        if size_divisor > 1:
            nw = math.ceil(nw / size_divisor) * size_divisor
            nh = math.ceil(nh / size_divisor) * size_divisor
            cv = resize_cubic(cv, nw, nh)
        outs.append(np.transpose(cv / np.float32(255), (2, 0, 1))[None].astype(np.float32))
    return outs[0], outs[1], ratio1, ratio2
