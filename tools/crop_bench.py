"""Timing of the device crop + bicubic resize (oetr_crop_resize) on a B200 against the reference's host round trip
(D2H -> cv2.resize INTER_CUBIC -> H2D, dloc/core/utils/utils.py:510-564) when cv2 is importable.   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from crop_cases import synthetic_image  # noqa: E402
from oetr_b200.dloc.core.utils import utils as U  # noqa: E402


def main():
    im1 = torch.from_numpy(synthetic_image(1, 1200, 1600, 1)).cuda()
    im2 = torch.from_numpy(synthetic_image(1, 1200, 1600, 2)).cuda()
    b1 = torch.tensor([[200.3, 100.9, 1400.2, 1000.5]]).cuda()
    b2 = torch.tensor([[50.0, 300.1, 900.7, 1150.0]]).cuda()
    for _ in range(3):
        l, r, _, _ = U.tensor_overlap_crop(im1, b1, im2, b2, "superpoint")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        l, r, _, _ = U.tensor_overlap_crop(im1, b1, im2, b2, "superpoint")
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n
    # the kernel alone (both crops in one launch), CUDA events
    jobs = [(im1[0], [200, 100, 1400, 1000], l.shape[3], l.shape[2], 3), (im2[0], [50, 300, 900, 1150], r.shape[3], r.shape[2], 3)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    U.crop_resize(jobs, im1.device)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        U.crop_resize(jobs, im1.device)
    e1.record()
    torch.cuda.synchronize()
    kern_ms = e0.elapsed_time(e1) / n
    out_bytes = (l.numel() + r.numel()) * 4
    in_bytes = (1200 * 900 + 850 * 850) * 4
    res = {"images": "1200x1600 gray, two crops per call", "out_shapes": [list(l.shape[2:]), list(r.shape[2:])],
           "call_ms_wall": wall * 1e3, "kernel_ms_incl_alloc": kern_ms, "algorithmic_bytes": in_bytes + out_bytes,
           "kernel_gbs": (in_bytes + out_bytes) / kern_ms / 1e6}
    try:
        import cv2
        def host(im, b):
            c = im[0, :, b[1]:b[3], b[0]:b[2]].permute(1, 2, 0).cpu().numpy() * 255
            return c
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            for im, b, o in ((im1, [200, 100, 1400, 1000], l), (im2, [50, 300, 900, 1150], r)):
                c = cv2.resize(host(im, b).astype("float32"), (o.shape[3], o.shape[2]), interpolation=cv2.INTER_CUBIC)
                t = torch.from_numpy(c / 255).float().to(im.device)
        torch.cuda.synchronize()
        res["reference_host_round_trip_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    except ImportError:
        res["reference_host_round_trip_ms"] = None
    print(json.dumps(res))


if __name__ == "__main__":
    main()
