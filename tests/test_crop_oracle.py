"""CPU suite of the post-box plumbing: the numpy oracle against cv2 itself (when importable) and against the
committed outputs of the reference's tensor_overlap_crop (tests/golden/crop_*.npz, made by make_crop_golden.py with cv2
4.13); gating logic; the C ABI's argument checks (no compute calls without a GPU).

Tolerance: cv2's optimised float32 bicubic and the oracle agree to float32 rounding (different summation order / FMA
contraction): 2e-4 on the 0..255 scale, i.e. 1e-6 on [0,1] images."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from crop_cases import CROP_CASES, synthetic_image
from oetr_b200 import cabi
from oracle import crop_oracle as co

TOL_255 = 2e-4


def test_resize_cubic_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (h, w, c, nw, nh) in [(37, 53, 1, 640, 447), (200, 150, 3, 480, 640), (64, 64, 1, 31, 17), (5, 7, 1, 40, 40),
                              (300, 400, 1, 304, 400), (100, 100, 3, 100, 100), (3, 2, 1, 9, 11), (1, 1, 1, 5, 4), (2, 9, 1, 9, 2)]:
        img = (rng.random((h, w, c)) * 255).astype(np.float32)
        img = img[:, :, 0] if c == 1 else img
        want = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_CUBIC)
        assert np.abs(co.resize_cubic(img, nw, nh) - want).max() < TOL_255, (h, w, c, nw, nh)


@pytest.mark.parametrize("name", sorted(CROP_CASES))
def test_tensor_overlap_crop_matches_reference_outputs(name):
    c, hw1, hw2, box1, box2, extractor, div, seed = CROP_CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", "crop_%s.npz" % name))
    left, right, r1, r2 = co.tensor_overlap_crop(synthetic_image(c, *hw1, seed), np.asarray([box1], np.float32),
                                                 synthetic_image(c, *hw2, seed + 100), np.asarray([box2], np.float32), extractor, div)
    assert left.shape == g["left"].shape and right.shape == g["right"].shape
    assert np.abs(left - g["left"]).max() < TOL_255 / 255 * 2 and np.abs(right - g["right"]).max() < TOL_255 / 255 * 2
    assert np.allclose(r1, g["ratio1"], rtol=0, atol=0) and np.allclose(r2, g["ratio2"], rtol=0, atol=0)


def test_overlap_gate_integer_logic():
    """evaluation.py:86-103: boxes are truncated to int; degenerate boxes (side <= 1) fall back to the full images; the
    pragueparks-val rule needs a scale ratio above 2."""
    assert co.overlap_gate([10.9, 10.9, 50.2, 60.7], [0, 0, 30, 30])
    assert not co.overlap_gate([10.9, 10.9, 11.99, 60.7], [0, 0, 30, 30])          # width int(11.99) - int(10.9) = 1
    assert not co.overlap_gate([0, 0, 100, 100], [5, 5, 5.5, 80])
    assert not co.overlap_gate([0, 0, 100, 100], [0, 0, 60, 60], "pragueparks-val")   # 100 // 60 = 1
    assert co.overlap_gate([0, 0, 100, 100], [0, 0, 30, 60], "pragueparks-val")       # 100 // 30 = 3
    from oetr_b200.dloc.core.utils import utils as U
    import torch
    for b0, b1, ds in (([10.9, 10.9, 50.2, 60.7], [0, 0, 30, 30], ""), ([10.9, 10.9, 11.99, 60.7], [0, 0, 30, 30], ""),
                       ([0, 0, 100, 100], [0, 0, 60, 60], "pragueparks-val"), ([0, 0, 100, 100], [0, 0, 30, 60], "pragueparks-val")):
        assert U.overlap_gate(torch.tensor([b0]), torch.tensor([b1]), ds) == co.overlap_gate(b0, b1, ds)
    for args in ((640, 480, 100, 50, "superpoint"), (640, 480, 50, 100, "superpoint"), (640, 480, 33, 77, "disk")):
        assert U.patch_resize(*args) == co.patch_resize(*args)


def test_crop_abi_argument_errors_do_not_need_a_gpu():
    lib = cabi.load_library()
    assert lib.oetr_crop_resize(None, 0, None) == 0
    assert lib.oetr_crop_resize(None, 2, None) == cabi.OETR_E_ARG
    from oetr_b200.dloc.core.utils.utils import _Job
    buf = (ctypes.c_float * 16)()
    addr = ctypes.addressof(buf)
    jobs = (_Job * 1)(_Job(addr, addr, 1, 4, 4, 2, 2, 2, 4, 8, 8, 0))              # empty crop (x1 == x0)
    assert lib.oetr_crop_resize(jobs, 1, None) == cabi.OETR_E_SHAPE and b"empty" in lib.oetr_crop_last_error()
    jobs = (_Job * 1)(_Job(None, addr, 1, 4, 4, 0, 0, 4, 4, 8, 8, 0))
    assert lib.oetr_crop_resize(jobs, 1, None) == cabi.OETR_E_ARG
    import torch
    from oetr_b200.dloc.core.utils import utils as U
    if not torch.cuda.is_available():
        with pytest.raises(cabi.OetrError):
            U.tensor_overlap_crop(torch.rand(1, 1, 8, 8), torch.tensor([[0., 0., 8., 8.]]), torch.rand(1, 1, 8, 8),
                                  torch.tensor([[0., 0., 8., 8.]]), "superpoint")
