"""Timing of the CUDA neck (oetr_neck_forward) on a B200 against eager PyTorch on the same GPU.
    python tools/neck_bench.py [--pairs 32] [--size 40] [--steps 20]
Prints one JSON line: pairs/s of the neck alone (both images of a pair), algorithmic TFLOP/s, per-kernel times."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oetr_b200 import weights  # noqa: E402
from oetr_b200.neck import NeckB200  # noqa: E402


def neck_flops(h, w):
    """multiply-add = 2, per image (20.3 GFLOP per 640 x 640 pair)"""
    t, o = h * w, (h // 2) * (w // 2)
    return 2.0 * (t * 1024 * 256 + o * 256 * (256 * 16 + 128 * 64 + 128 * 256) + o * 512 * 256)


def timed(fn, steps, warmup=3):
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


def torch_neck(W, dtype, channels_last):
    import torch.nn.functional as F
    Wt = {k: torch.from_numpy(v).cuda() for k, v in W.items()}
    cw = {k: (v.to(dtype).contiguous(memory_format=torch.channels_last) if (channels_last and v.dim() == 4) else v.to(dtype))
          for k, v in Wt.items()}

    def run(x):
        x = x.to(dtype)
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        p = F.conv2d(x, cw["input_proj.weight"], cw["input_proj.bias"])
        p = F.layer_norm(p.permute(0, 2, 3, 1), (256,), cw["patchmerging.norm.weight"], cw["patchmerging.norm.bias"]).permute(0, 3, 1, 2)
        outs = [F.conv2d(p, cw["patchmerging.reductions.%d.weight" % i], cw["patchmerging.reductions.%d.bias" % i], stride=2,
                         padding=(k - 2) // 2) for i, k in enumerate((4, 8, 16))]
        return F.conv2d(torch.cat(outs, 1), cw["input_proj2.weight"], cw["input_proj2.bias"]).float()
    return run


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--size", type=int, default=40)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-torch", action="store_true")
    a = ap.parse_args()
    n, h, w = 2 * a.pairs, a.size, a.size
    W = weights.synthetic_neck_weights(0)
    xs = [torch.from_numpy(weights.synthetic_backbone_features(n, h, w, seed=50 + i)).cuda() for i in range(2)]   # 2 x 420 MB > L2
    neck = NeckB200(W)
    out = torch.empty(n, 256, h // 2, w // 2, device="cuda")
    it = [0]

    def step():
        neck.forward(xs[it[0] & 1], out=out)
        it[0] += 1
    ms = timed(step, a.steps)
    fl = neck_flops(h, w) * n
    res = {"metric": "neck image-pairs/sec", "pairs": a.pairs, "backbone_hw": [h, w], "ms_per_step": ms,
           "value": a.pairs / ms * 1e3, "unit": "pairs/s", "algorithmic_tflops": fl / ms / 1e9,
           "gflop_per_pair": fl / a.pairs / 1e9}
    if not a.no_torch:
        ref32 = torch_neck(W, torch.float32, False)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        want = ref32(xs[0])
        got = neck.forward(xs[0])
        torch.cuda.synchronize()
        res["max_err_vs_torch_fp32_over_std"] = float((got - want).abs().max() / want.std())
        res["torch_fp32_ms"] = timed(lambda: ref32(xs[0]), 5, 2)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        res["torch_tf32_ms"] = timed(lambda: ref32(xs[0]), 10, 2)
        refbf = torch_neck(W, torch.bfloat16, True)
        res["torch_bf16_channels_last_ms"] = timed(lambda: refbf(xs[0]), 10, 3)
        res["torch_bf16_max_err_over_std"] = float((refbf(xs[0]) - want).abs().max() / want.std())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
