"""Parity of the device-side crop + bicubic resize (oetr_crop_resize through the mirror of the reference's
tensor_overlap_crop) against the CPU oracle and the committed outputs of the reference function (cv2 4.13).
Tolerance: float32 rounding of the interpolation sums, 1e-6 on [0,1] images (2e-4 on cv2's 0..255 scale)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from crop_cases import CROP_CASES, synthetic_image
from oracle import crop_oracle as co

pytestmark = pytest.mark.gpu

from oetr_b200.dloc.core.utils import utils as U  # noqa: E402

TOL = 1.6e-6


@pytest.mark.parametrize("name", sorted(CROP_CASES))
def test_tensor_overlap_crop_matches_reference_and_oracle(name):
    c, hw1, hw2, box1, box2, extractor, div, seed = CROP_CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", "crop_%s.npz" % name))
    im1, im2 = synthetic_image(c, *hw1, seed), synthetic_image(c, *hw2, seed + 100)
    left, right, r1, r2 = U.tensor_overlap_crop(torch.from_numpy(im1).cuda(), torch.tensor([box1]).cuda(),
                                                torch.from_numpy(im2).cuda(), torch.tensor([box2]).cuda(), extractor, div)
    assert left.is_cuda and left.dtype == torch.float32
    wl, wr, q1, q2 = co.tensor_overlap_crop(im1, np.asarray([box1], np.float32), im2, np.asarray([box2], np.float32), extractor, div)
    for got, want, gold in ((left, wl, g["left"]), (right, wr, g["right"])):
        got = got.cpu().numpy()
        assert got.shape == gold.shape
        assert np.abs(got - want).max() < TOL and np.abs(got - gold).max() < TOL
    assert r1 == q1 and r2 == q2 and np.allclose(r1, g["ratio1"], atol=0) and np.allclose(r2, g["ratio2"], atol=0)


def test_full_size_properties():
    """1200 x 1600 images (the evaluation's resize target): identity box + same size = exact copy; a resize and its
    oracle agree on a strided sample of rows; deterministic."""
    im = torch.from_numpy(synthetic_image(1, 1200, 1600, 9)).cuda()
    full = torch.tensor([[0.0, 0.0, 1600.0, 1200.0]]).cuda()
    l, r, r1, r2 = U.tensor_overlap_crop(im, full, im, full, "superpoint")
    want_id = im.cpu().numpy() * np.float32(255) / np.float32(255)          # numpy's true division, like the reference (torch multiplies by 1/255)
    assert np.array_equal(l.cpu().numpy(), want_id) and r1 == [[1.0, 1.0]]
    box = torch.tensor([[100.7, 50.2, 1300.1, 1100.9]]).cuda()
    a, _, ra, _ = U.tensor_overlap_crop(im, box, im, full, "superpoint")
    b, _, _, _ = U.tensor_overlap_crop(im, box, im, full, "superpoint")
    assert torch.equal(a, b) and a.shape[2] == 1200 and a.shape[3] == int(1200 / 1050 * 1200)
    want, _, rw, _ = co.tensor_overlap_crop(im.cpu().numpy(), box.cpu().numpy(), im.cpu().numpy(), full.cpu().numpy(), "superpoint")
    assert ra == rw and np.abs(a.cpu().numpy() - want).max() < TOL


def test_batched_jobs_and_errors():
    ims = [torch.from_numpy(synthetic_image(3, 60 + 7 * i, 80 + 5 * i, 20 + i))[0].cuda() for i in range(40)]     # > 32 jobs: two launches
    jobs = [(im, (3, 2, im.shape[2] - 4, im.shape[1] - 1), 50 + i, 40 + 2 * i, U.MUL255 | U.DIV255) for i, im in enumerate(ims)]
    outs = U.crop_resize(jobs, ims[0].device)
    for i in (0, 17, 39):
        im, box, nw, nh, _ = jobs[i]
        cv = np.transpose(im.cpu().numpy()[:, box[1]:box[3], box[0]:box[2]], (1, 2, 0)) * np.float32(255)
        want = np.transpose(co.resize_cubic(cv, nw, nh) / np.float32(255), (2, 0, 1))
        assert np.abs(outs[i].cpu().numpy() - want).max() < TOL, i
    from oetr_b200 import cabi
    with pytest.raises(cabi.OetrError):
        U.crop_resize([(ims[0], (10, 10, 10, 30), 8, 8, 0)], ims[0].device)
