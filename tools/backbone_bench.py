"""Backbone execution modes on a B200: images/s of the ResNet-50 trunk (to layer3) per mode, the feature
error against the reference's eager fp32 execution, and the box error each mode causes through the CUDA neck + hot path.
    python tools/backbone_bench.py [--pairs 32] [--size 640]
Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import oetr_b200  # noqa: E402
from pipeline_sample import synthetic_state_dict  # noqa: E402


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = oetr_b200.build_detectors(oetr_b200.get_cfg_defaults().OETR)
    net.load_state_dict(synthetic_state_dict(net))
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(0)
    img1 = torch.rand((a.pairs, a.size, a.size, 3), generator=g).cuda()
    img2 = torch.rand((a.pairs, a.size, a.size, 3), generator=g).cuda()
    res = {"metric": "backbone images/sec (ResNet-50 to layer3)", "images": 2 * a.pairs, "image_size": a.size, "modes": {}}
    with torch.no_grad():
        net.backbone.set_execution_mode("eager")
        f_ref = net.backbone(img1)
        b_ref = net.forward_dummy(img1, img2)
        for mode, graphs in (("eager", False), ("channels_last", False), ("channels_last", True), ("tf32", True), ("bf16", False), ("bf16", True)):
            net.backbone.set_execution_mode(mode, graphs=graphs)
            f = net.backbone(img1)
            b = net.forward_dummy(img1, img2)
            ms = timed(lambda: (net.backbone(img1), net.backbone(img2)))
            res["modes"]["%s%s" % (mode, "+graphs" if graphs else "")] = {
                "ms_per_batch": ms, "images_per_s": 2 * a.pairs / ms * 1e3,
                "feature_rel_err": float((f - f_ref).abs().max() / f_ref.abs().max()),
                "box_err_over_side": float(max((b[0] - b_ref[0]).abs().max(), (b[1] - b_ref[1]).abs().max()) / a.size)}
        net.backbone.set_execution_mode("eager")
        x = torch.cat([net.backbone(img1), net.backbone(img2)], dim=0)
        res["neck_ms"] = timed(lambda: net.neck_path.forward(x))
        f = net.neck_path.forward(x)
        f1, f2 = f[: a.pairs].contiguous(), f[a.pairs:].contiguous()
        res["hot_path_ms_one_batch_in_flight"] = timed(lambda: net.feature_correlation_and_regression(f1, f2, (a.size, a.size), (a.size, a.size)))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
