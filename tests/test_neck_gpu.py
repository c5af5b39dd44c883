"""Parity of the CUDA neck (oetr_neck_forward through ctypes) against the CPU oracle and the committed outputs of the
reference modules.

Tolerance: the neck runs single fp16 operands with fp32 accumulation.  The oracle's rounding model predicts a feature error
of 5e-4 relative rms / 2.3e-3 of the feature std at the maximum, and a box error of 1.1e-4 of the image side after the hot
path (bar 1e-3).  Bars used here: max |error| < 4e-3 * std, rms error < 1e-3 * std, boxes < 1e-3.
"""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from neck_cases import NECK_CASES
from oracle import neck_oracle as nk
from oracle import oetr_oracle as orc

pytestmark = pytest.mark.gpu

import oetr_b200  # noqa: E402
from oetr_b200 import weights  # noqa: E402
from oetr_b200.neck import NeckB200  # noqa: E402

MAX_TOL, RMS_TOL = 4e-3, 1e-3


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", "neck_%s.npz" % name))


def _check(got, want, std, what):
    err = np.abs(np.asarray(got, np.float64) - want)
    assert err.max() < MAX_TOL * std, (what, "max", err.max() / std)
    assert np.sqrt((err ** 2).mean()) < RMS_TOL * std, (what, "rms", np.sqrt((err ** 2).mean()) / std)


@pytest.mark.parametrize("name", sorted(NECK_CASES))
def test_neck_matches_reference_outputs(name):
    n, h, w, wseed, fseed, gain = NECK_CASES[name]
    W = weights.synthetic_neck_weights(wseed, gain=gain)
    x = weights.synthetic_backbone_features(n, h, w, seed=fseed)
    neck = NeckB200(W)
    f = neck.forward(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert neck.last_launch_count == 3
    f = f.cpu().numpy()
    g = _golden(name)
    assert f.shape == (n, 256, h // 2, w // 2) and np.isfinite(f).all()
    _check(f[:, ::4], g["feat_c4"], float(g["std"]), name)
    assert np.allclose(f.astype(np.float64).sum(axis=(1, 2, 3)), g["sums"], rtol=0, atol=2e-3 * float(g["abs_sums"].max()))
    neck.close()


def test_neck_matches_oracle_all_channels():
    n, h, w = 2, 14, 10
    W = weights.synthetic_neck_weights(5)
    x = weights.synthetic_backbone_features(n, h, w, seed=21)
    want = nk.neck(W, x)
    neck = NeckB200(W)
    got = neck.forward(torch.from_numpy(x).cuda()).cpu().numpy()
    _check(got, want, float(want.std()), "14x10")
    # the rounding model of the kernels (fp16 operands, wide accumulation) explains most of the difference
    model = nk.neck(W, x, rnd_proj=orc.round_fp16, rnd_conv=orc.round_fp16, rnd_proj2=orc.round_fp16)
    assert np.abs(got - model).max() < np.abs(got - want).max()
    neck.close()


def test_neck_full_batch_properties():
    """BASELINE config 2's image count (64 images of 40x40; 220 conv tiles x 2 split-K parts): deterministic; the two images
    of the golden case, embedded in the batch, match the reference's outputs; every image equals its single-image run up
    to the accumulation order (another tiling and 16 parts instead of 2; the tensor core's fp32 accumulator truncates, so a
    chain of 2 816 accumulations differs from one of 352 by up to ~6e-4 of the feature std -- measured 5.4e-4)."""
    W = weights.synthetic_neck_weights(0)
    xh = weights.synthetic_backbone_features(64, 40, 40, seed=31)
    n_g, h_g, w_g, wseed, fseed, _ = NECK_CASES["40x40"]
    assert wseed == 0 and (h_g, w_g) == (40, 40)
    xh[7:7 + n_g] = weights.synthetic_backbone_features(n_g, 40, 40, seed=fseed)
    x = torch.from_numpy(xh).cuda()
    neck = NeckB200(W)
    a = neck.forward(x).clone()
    b = neck.forward(x)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    std = float(a.std())
    assert 0.2 < std < 0.5
    for i in (0, 5, 63):
        one = neck.forward(x[i:i + 1].contiguous())
        assert float((one[0] - a[i]).abs().max()) < 1.5e-3 * std, i
    g = _golden("40x40")
    _check(a[7:7 + n_g, ::4].cpu().numpy(), g["feat_c4"], float(g["std"]), "40x40 inside the batch of 64")
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0)).cuda()
    c = neck.forward(x[perm].contiguous())
    assert float((c - a[perm]).abs().max()) < 1.5e-3 * std        # images move to other tile rows / tiles
    neck.close()


def test_neck_then_hot_path_boxes():
    """images' backbone features -> CUDA neck -> CUDA hot path against neck oracle -> hot-path oracle: < 1e-3 of the side."""
    n, h, w = 2, 40, 40
    NW, HW = weights.synthetic_neck_weights(0), weights.synthetic_hot_path_weights(0)
    x1 = weights.synthetic_backbone_features(n, h, w, seed=41, tag="a")
    x2 = weights.synthetic_backbone_features(n, h, w, seed=41, tag="b")
    want = orc.hot_path(HW, nk.neck(NW, x1), nk.neck(NW, x2), (640, 640), (640, 640), clamp=False)
    neck = NeckB200(NW)
    hot = oetr_b200.OverlapHotPath(HW, precision="fp16")
    f = neck.forward(torch.from_numpy(np.concatenate([x1, x2])).cuda())          # both image sets in one call
    b1, b2 = hot.forward(f[:n].contiguous(), f[n:].contiguous(), (640, 640), (640, 640), clamp=False)
    torch.cuda.synchronize()
    e1 = np.abs(b1.cpu().numpy() - want["box1_raw"]).max() / 640
    e2 = np.abs(b2.cpu().numpy() - want["box2_raw"]).max() / 640
    assert e1 < 1e-3 and e2 < 1e-3, (e1, e2)
    neck.close(); hot.close()


def test_neck_errors():
    W = weights.synthetic_neck_weights(0)
    neck = NeckB200(W)
    with pytest.raises(ValueError):
        neck.forward(torch.zeros(1, 256, 40, 40, device="cuda"))
    with pytest.raises(oetr_b200.cabi.OetrError):
        neck.forward(torch.zeros(1, 1024, 1, 40, device="cuda"))
    neck.close()
