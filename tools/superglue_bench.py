"""Timing of SuperGlue's two CUDA operators on a B200 against eager PyTorch restatements of the reference's functions
(superglue.py:86-90, :143-184) on the same GPU.   python tools/superglue_bench.py [--kpts 2048]   -> one JSON line"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oetr_b200 import superglue as sg  # noqa: E402


def timed(fn, steps=20, warmup=3):
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


def torch_attention(q, k, v):
    s = torch.einsum("bdhn,bdhm->bhnm", q, k) / 8.0
    return torch.einsum("bhnm,bdhm->bdhn", torch.softmax(s, dim=-1), v).contiguous()


def torch_transport(scores, alpha, iters):
    b, m, n = scores.shape
    a = scores.new_tensor(alpha)
    Z = torch.cat([torch.cat([scores, a.expand(b, m, 1)], -1), torch.cat([a.expand(b, 1, n), a.expand(b, 1, 1)], -1)], 1)
    norm = -torch.log(scores.new_tensor(float(m + n)))
    log_mu = torch.cat([norm.expand(m), torch.log(scores.new_tensor(float(n)))[None] + norm])[None]
    log_nu = torch.cat([norm.expand(n), torch.log(scores.new_tensor(float(m)))[None] + norm])[None]
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1) - norm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kpts", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=100)
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    n = a.kpts
    q, k, v = (torch.randn(1, 64, 4, n, generator=g).cuda() for _ in range(3))
    s = (torch.randn(1, n, n, generator=g) * 2).cuda()
    res = {"keypoints": n, "sinkhorn_iterations": a.iters}
    res["attention_ms"] = timed(lambda: sg.attention(q, k, v))
    res["attention_fp32_kernel_ms"] = timed(lambda: sg.attention(q, k, v, mode="fp32"))
    res["attention_torch_ms"] = timed(lambda: torch_attention(q, k, v))
    res["attention_max_diff"] = float((sg.attention(q, k, v) - torch_attention(q, k, v)).abs().max())
    res["attention_gflop"] = 4 * 2 * 2.0 * n * n * 64 / 1e9
    res["attention_tflops"] = res["attention_gflop"] / res["attention_ms"]
    res["transport_ms"] = timed(lambda: sg.log_optimal_transport(s, 1.0, a.iters), steps=5)
    res["transport_torch_ms"] = timed(lambda: torch_transport(s, 1.0, a.iters), steps=3, warmup=1)
    res["transport_max_diff"] = float((sg.log_optimal_transport(s, 1.0, a.iters) - torch_transport(s, 1.0, a.iters)).abs().max())
    # one Sinkhorn half-iteration reads the n x n scores once: effective bandwidth (L2-resident at this size)
    res["transport_gbs"] = 2 * a.iters * n * n * 4 / res["transport_ms"] / 1e6
    # the whole SuperGlue forward (mirror with the CUDA operators) against the same module with PyTorch operators
    torch.manual_seed(0)
    model = sg.SuperGlue({"sinkhorn_iterations": a.iters}).cuda().eval()
    data = {"image0": torch.zeros(1, 1, 1200, 1600), "image1": torch.zeros(1, 1, 1200, 1600),
            "keypoints0": (torch.rand(1, n, 2, generator=g) * torch.tensor([1600.0, 1200.0])).cuda(),
            "keypoints1": (torch.rand(1, n, 2, generator=g) * torch.tensor([1600.0, 1200.0])).cuda(),
            "scores0": torch.rand(1, n, generator=g).cuda(), "scores1": torch.rand(1, n, generator=g).cuda(),
            "descriptors0": torch.nn.functional.normalize(torch.randn(1, 256, n, generator=g), dim=1).cuda(),
            "descriptors1": torch.nn.functional.normalize(torch.randn(1, 256, n, generator=g), dim=1).cuda()}
    out = model(data)
    res["model_ms"] = timed(lambda: model(data), steps=5, warmup=2)
    att, ot = sg.attention, sg.log_optimal_transport
    sg.attention = lambda q_, k_, v_, mode="tensor": torch_attention(q_, k_, v_)
    sg.log_optimal_transport = lambda s_, alpha, iters: torch_transport(s_, float(alpha), iters)
    ref = model(data)
    res["model_torch_ops_ms"] = timed(lambda: model(data), steps=3, warmup=1)
    sg.attention, sg.log_optimal_transport = att, ot
    res["model_matches_equal"] = bool(torch.equal(out["matches0"], ref["matches0"]))
    res["model_matches"] = int((out["matches0"] > -1).sum())
    res["model_scores_max_diff"] = float((out["scores"] - ref["scores"]).abs().max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
