"""Parity of SuperGlue's CUDA operators (oetr_sg_attention, oetr_sg_optimal_transport through ctypes) and of the SuperGlue
mirror built on them against the CPU oracle and the committed outputs of the reference module.
Tolerances (against an fp32 reference / an fp64 oracle): attention 2e-5 absolute on O(1) outputs for the fp32 kernel and 6e-5
for the tcgen05 kernel (3-term split fp16 operands: 2^-22 per product, amplified by the softmax), transport 2e-4 on
log-scores, full model 2e-3 on log-scores with identical matches."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import superglue_oracle as so

pytestmark = pytest.mark.gpu

from oetr_b200 import superglue as sg  # noqa: E402

OPS = np.load(os.path.join(ROOT, "tests", "golden", "superglue_ops.npz"))
WEIGHT_PATHS = [os.path.join(ROOT, "oracle", "_ref", "weights", "superglue_outdoor.pth"),
                "/root/reference/third_party/SuperGluePretrainedNetwork/models/weights/superglue_outdoor.pth"]


ATT_TOL = {"fp32": 2e-5, "tensor": 6e-5}


@pytest.mark.parametrize("mode", ["tensor", "fp32"])
@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_attention_matches_reference_outputs(name, mode):
    q, k, v, want = (torch.from_numpy(OPS["att_%s_%s" % (name, t)]) for t in ("q", "k", "v", "out"))
    got = sg.attention(q.cuda(), k.cuda(), v.cuda(), mode=mode).cpu()
    assert got.shape == want.shape and float((got - want).abs().max()) < ATT_TOL[mode], float((got - want).abs().max())


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_optimal_transport_matches_reference_outputs(name):
    s, want, it = torch.from_numpy(OPS["ot_%s_in" % name]), OPS["ot_%s_out" % name], int(OPS["ot_%s_iters" % name])
    got = sg.log_optimal_transport(s.cuda(), torch.tensor(2.3), it).cpu().numpy()
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-4
    for b in range(s.shape[0]):
        assert np.abs(got[b] - so.log_optimal_transport(s[b].numpy().astype(np.float64), 2.3, it)).max() < 2e-4


def test_large_problem_properties():
    """2048 x 2048 keypoints (the evaluation's cap): attention rows are convex combinations of V; the transport plan's
    marginals are the prescribed ones after convergence; attention agrees with the oracle on a sample of queries."""
    g = torch.Generator().manual_seed(1)
    n = m = 2048
    q, k, v = (torch.randn(1, 64, 4, s, generator=g).cuda() for s in (n, m, m))
    idx = [0, 1, 63, 64, 127, 128, 1000, 2047]
    want = so.attention(q[0][:, :, idx].double().cpu().numpy(), k[0].double().cpu().numpy(), v[0].double().cpu().numpy())
    for mode in ("tensor", "fp32"):
        o = sg.attention(q, k, v, mode=mode)
        assert float(o.max()) <= float(v.max()) + 1e-4 and float(o.min()) >= float(v.min()) - 1e-4
        assert np.abs(o[0][:, :, idx].cpu().numpy() - want).max() < ATT_TOL[mode], mode
        assert torch.equal(o, sg.attention(q, k, v, mode=mode))
    # key split over CTA clusters: 128 queries x 600 keys -> 5 chunks over 4 CTAs (one peer gets none); 256 x 1024 -> 2 x 4 CTAs
    for nq, nk in ((128, 600), (256, 1024), (100, 129)):
        qs, ks, vs = q[:, :, :, :nq].contiguous(), k[:, :, :, :nk].contiguous(), v[:, :, :, :nk].contiguous()
        wants = so.attention(qs[0].double().cpu().numpy(), ks[0].double().cpu().numpy(), vs[0].double().cpu().numpy())
        assert np.abs(sg.attention(qs, ks, vs)[0].cpu().numpy() - wants).max() < ATT_TOL["tensor"], (nq, nk)
    # ragged sizes: 130 queries (two query tiles, the second almost empty) x 257 keys (three key chunks, the last with one key)
    q2, k2, v2 = q[:, :, :, :130].contiguous(), k[:, :, :, :257].contiguous(), v[:, :, :, :257].contiguous()
    want2 = so.attention(q2[0].double().cpu().numpy(), k2[0].double().cpu().numpy(), v2[0].double().cpu().numpy())
    for mode in ("tensor", "fp32"):
        assert np.abs(sg.attention(q2, k2, v2, mode=mode)[0].cpu().numpy() - want2).max() < ATT_TOL[mode], mode
    s = torch.randn(1, m, n, generator=g).cuda() * 2
    Z = sg.log_optimal_transport(s, torch.tensor(1.0), 100)
    P = (Z.double() - np.log(m + n)).exp()                   # undo the (m + n) scaling of :183
    mu = 1.0 / (m + n)
    assert float((P[0, :, :n].sum(0) / mu - 1.0).abs().max()) < 1e-4      # the last half-iteration fixes the column marginals
    assert float((P[0, :m].sum(1) / mu - 1.0).abs().max()) < 1e-2         # rows: converged
    assert torch.equal(Z, sg.log_optimal_transport(s, torch.tensor(1.0), 100))


def _weights():
    for p in WEIGHT_PATHS:
        if os.path.exists(p):
            return p
    return None


def test_superglue_mirror_with_synthetic_weights_matches_oracle():
    """The whole mirror (PyTorch 1x1 convolutions + CUDA attention + CUDA transport) against the oracle with the same
    deterministic weights (no trained weights needed on the GPU box)."""
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = sg.SuperGlue({"sinkhorn_iterations": 30}).eval()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 3:
                p.mul_(1.5)
        for name, b in model.named_buffers():
            if name.endswith("running_var"):
                b.uniform_(0.5, 1.5)
            elif name.endswith("running_mean"):
                b.uniform_(-0.2, 0.2)
    W = {k: v.numpy() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    n0, n1 = 150, 97
    data = {"image0": torch.zeros(1, 1, 480, 640), "image1": torch.zeros(1, 1, 400, 300),
            "keypoints0": torch.rand(1, n0, 2, generator=g) * torch.tensor([640.0, 480.0]),
            "keypoints1": torch.rand(1, n1, 2, generator=g) * torch.tensor([300.0, 400.0]),
            "scores0": torch.rand(1, n0, generator=g), "scores1": torch.rand(1, n1, generator=g),
            "descriptors0": torch.nn.functional.normalize(torch.randn(1, 256, n0, generator=g), dim=1),
            "descriptors1": torch.nn.functional.normalize(torch.randn(1, 256, n1, generator=g), dim=1)}
    want = so.superglue(W, data["keypoints0"][0].numpy(), data["keypoints1"][0].numpy(), data["scores0"][0].numpy(),
                        data["scores1"][0].numpy(), data["descriptors0"][0].numpy(), data["descriptors1"][0].numpy(),
                        (480, 640), (400, 300), iters=30)
    model = model.cuda()
    out = model({k: v.cuda() for k, v in data.items()})
    assert np.abs(out["scores"][0].cpu().numpy() - want["scores"]).max() < 2e-3
    assert np.array_equal(out["matches0"][0].cpu().numpy(), want["matches0"])
    assert np.array_equal(out["matches1"][0].cpu().numpy(), want["matches1"])
    empty = dict(data, keypoints0=torch.zeros(1, 0, 2))
    assert model({k: v.cuda() for k, v in empty.items()})["matches1"].shape == (1, n1)


@pytest.mark.skipif(_weights() is None, reason="superglue_outdoor.pth not shipped (oracle/_ref/weights is filled by build() in the build container)")
def test_superglue_mirror_with_the_in_tree_weights_matches_reference_outputs():
    g = np.load(os.path.join(ROOT, "tests", "golden", "superglue_pair1.npz"))
    model = sg.SuperGlue({"weights": _weights(), "sinkhorn_iterations": 50}).cuda().eval()
    data = {k: torch.from_numpy(g[k]).cuda() for k in ("keypoints0", "keypoints1", "scores0", "scores1", "descriptors0", "descriptors1")}
    data["image0"], data["image1"] = torch.zeros(*g["shape0"]), torch.zeros(*g["shape1"])
    out = model(data)
    assert np.abs(out["scores"][0].cpu().numpy() - g["scores"][0]).max() < 2e-3
    assert np.array_equal(out["matches0"][0].cpu().numpy(), g["matches0"][0])
    assert np.array_equal(out["matches1"][0].cpu().numpy(), g["matches1"][0])
    assert np.abs(out["matching_scores0"][0].cpu().numpy() - g["matching_scores0"][0]).max() < 1e-3
