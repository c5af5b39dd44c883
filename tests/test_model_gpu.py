"""The drop-in boundary on a B200: the mirror of the reference model class (images -> PyTorch backbone + neck -> CUDA hot
path -> boxes) and the dloc overlap plugin, checked against the CPU oracle run on the SAME backbone features."""
import numpy as np
import pytest
import torch

from oracle import oetr_oracle as orc

pytestmark = pytest.mark.gpu

import oetr_b200  # noqa: E402
from oetr_b200.dloc.core import overlap_features, overlaps  # noqa: E402
from oetr_b200.dloc.core.utils.base_model import dynamic_load  # noqa: E402


def _model(precision="fp16"):
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return oetr_b200.build_detectors(oetr_b200.get_cfg_defaults().OETR, precision=precision).cuda().eval()


def _oracle_boxes(model, img1, img2, clamp):
    with torch.no_grad():
        f1, f2 = model.feature_extraction(img1, img2)
    W = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    hw1, hw2 = tuple(img1.shape[1:3]), tuple(img2.shape[1:3])
    o = orc.hot_path(W, f1.cpu().numpy(), f2.cpu().numpy(), hw1, hw2, clamp=clamp)
    key = ("box1", "box2") if clamp else ("box1_raw", "box2_raw")
    return o[key[0]], o[key[1]], hw1, hw2


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_forward_dummy_and_forward_match_the_oracle_on_backbone_features(precision):
    model = _model(precision)
    g = torch.Generator().manual_seed(1)
    # two geometries: small ragged maps (per-image encoder tiles) and 640x640 (flat tiles)
    for shape1, shape2 in (((2, 320, 416, 3), (2, 384, 288, 3)), ((2, 640, 640, 3), (2, 640, 640, 3))):
        img1, img2 = torch.rand(shape1, generator=g).cuda(), torch.rand(shape2, generator=g).cuda()
        b1, b2 = model.forward_dummy(img1, img2)                      # reference src/model.py:229-252
        assert b1.shape == (2, 4) and b2.shape == (2, 4) and b1.is_cuda and b1.dtype == torch.float32
        w1, w2, hw1, hw2 = _oracle_boxes(model, img1, img2, clamp=True)
        tol = 1e-3 if precision == "fp16" else 2e-5
        assert np.abs(b1.cpu().numpy() - w1).max() / max(hw1) < tol
        assert np.abs(b2.cpu().numpy() - w2).max() / max(hw2) < tol
        out = model({"image1": img1, "image2": img2})                 # training-signature entry, unclamped (:193-211)
        r1, r2, _, _ = _oracle_boxes(model, img1, img2, clamp=False)
        assert np.abs(out["pred_bbox1"].cpu().numpy() - r1).max() / max(hw1) < tol
        assert np.abs(out["pred_bbox2"].cpu().numpy() - r2).max() / max(hw2) < tol
    with pytest.raises(ValueError):
        model.forward_dummy(img1, img2, mask1=torch.ones(2, 20, 20))            # both masks or neither
    # masks at feature-map resolution (reference model.py:229-250): against the oracle with the same masks
    from oetr_b200 import weights
    m1 = torch.from_numpy(weights.synthetic_mask(2, 20, 20, tag="mask1")).cuda()
    m2 = torch.from_numpy(weights.synthetic_mask(2, 20, 20, tag="mask2")).cuda()
    b1, b2 = model.forward_dummy(img1, img2, mask1=m1, mask2=m2)
    with torch.no_grad():
        f1, f2 = model.feature_extraction(img1, img2)
    W = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    o = orc.hot_path(W, f1.cpu().numpy(), f2.cpu().numpy(), (640, 640), (640, 640), mask1=m1.cpu().numpy(),
                     mask2=m2.cpu().numpy())
    assert np.abs(b1.cpu().numpy() - o["box1"]).max() / 640 < tol and np.abs(b2.cpu().numpy() - o["box2"]).max() / 640 < tol


def test_dloc_plugin_runs_like_the_reference_plugin(tmp_path):
    """reference dloc/core/overlaps/oetr.py:15-46 and its caller evaluation.py:77-92: batch of one, tuple of two
    [1,4] fp32 tensors on the input device, usable as `bbox * scales` and `.int()`."""
    model = _model()
    (tmp_path / "oetr").mkdir()
    torch.save(model.state_dict(), tmp_path / "oetr" / "x.pth")
    conf = dict(overlap_features.confs["oetr"]["model"], weights="oetr/x.pth")
    plug = dynamic_load(overlaps, "oetr")(conf, tmp_path).cuda().eval()
    g = torch.Generator().manual_seed(2)
    img0, img1 = torch.rand((1, 480, 640, 3), generator=g).cuda(), torch.rand((1, 640, 480, 3), generator=g).cuda()
    with torch.no_grad():
        box0, box1 = plug({"image0": img0, "image1": img1})
    assert isinstance(box0, torch.Tensor) and box0.shape == (1, 4) and box0.is_cuda and box0.dtype == torch.float32
    w0, w1, hw0, hw1 = _oracle_boxes(plug.net, img0, img1, clamp=True)
    assert np.abs(box0.cpu().numpy() - w0).max() / max(hw0) < 1e-3
    assert np.abs(box1.cpu().numpy() - w1).max() / max(hw1) < 1e-3
    scales = torch.tensor([1.5, 2.0, 1.5, 2.0], device=box0.device)
    assert (box0 * scales)[0].int().shape == (4,)
