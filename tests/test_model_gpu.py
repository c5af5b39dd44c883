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
        f1, f2 = model.feature_extraction(img1, img2)[:2]
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
        f1, f2 = model.feature_extraction(img1, img2)[:2]
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


def test_stage_wise_signatures_of_the_reference():
    """The reference's own stage methods (src/model.py:109-191, transformer.py:313-383) on the mirror: the 8-tuple of
    feature_extraction, feature_correlation -> (hs1, hs2, memory1, memory2), center_estimation, size_regression, and the
    stand-alone QueryTransformer.forward -- the sequence forward_dummy (model.py:236-250) runs, against the oracle."""
    from oetr_b200.model import QueryTransformer
    model = _model("fp16")
    g = torch.Generator().manual_seed(3)
    img1, img2 = torch.rand((2, 320, 416, 3), generator=g).cuda(), torch.rand((2, 384, 288, 3), generator=g).cuda()
    model.h1, model.w1 = img1.shape[1:3]
    model.h2, model.w2 = img2.shape[1:3]
    with torch.no_grad():
        feat1, feat2, pos1, pos2, hf1, wf1, hf2, wf2 = model.feature_extraction(img1, img2)
    assert pos1.shape == (1, 256, hf1, wf1) and (hf1, wf1, hf2, wf2) == (10, 13, 12, 9)
    hs1, hs2, mem1, mem2 = model.feature_correlation(feat1, feat2, pos1, pos2, None, None)
    assert hs1.shape == (2, 1, 256) and mem1.shape == (2, hf1 * wf1, 256) and mem2.shape == (2, hf2 * wf2, 256)
    cxy1, cxy2 = model.center_estimation(hs1, hs2, mem1, mem2, hf1, wf1, hf2, wf2, None, None)
    tlbr1, tlbr2 = model.size_regression(hs1, hs2)
    W = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    o = orc.hot_path(W, feat1.cpu().numpy(), feat2.cpu().numpy(), (320, 416), (384, 288), clamp=False)
    rel = lambda a, b: float(np.abs(a.cpu().numpy() - b).max() / np.abs(b).max())
    assert rel(hs1[:, 0], o["hs1"]) < 3e-3 and rel(mem2, o["memory2"]) < 3e-3
    assert np.abs(cxy1.cpu().numpy() - o["cxy1"]).max() / 416 < 1e-3 and np.abs(cxy2.cpu().numpy() - o["cxy2"]).max() / 384 < 1e-3
    assert rel(tlbr1, o["tlbr1"]) < 3e-3 and rel(tlbr2, o["tlbr2"]) < 3e-3
    # the head takes ANY hs / memory (here: the oracle's), like the reference's method
    c1, _ = model.center_estimation(torch.from_numpy(o["hs1"][:, None]).float().cuda(), hs2,
                                    torch.from_numpy(o["memory1"]).float().cuda(), mem2, hf1, wf1, hf2, wf2)
    assert np.abs(c1.cpu().numpy() - o["cxy1"]).max() / 416 < 2e-5
    # stand-alone transformer with the reference's forward signature
    qt = QueryTransformer(256, 8, 4).cuda().eval()
    qt.load_state_dict(model.transformer.state_dict())
    h1, h2, m1, m2 = qt(feat1, feat2, model.query_embed1.weight, model.query_embed2.weight, pos1, pos2, None, None)
    assert rel(h1[:, 0], o["hs1"]) < 3e-3 and rel(m1, o["memory1"]) < 3e-3
    # in-place weight updates are seen (the CUDA handle snapshots the weights; it is rebuilt on a version change)
    b_before, _ = model.forward_dummy(img1, img2)
    with torch.no_grad():
        model.tlbr_reg[2].bias.add_(0.5)
    b_after, _ = model.forward_dummy(img1, img2)
    assert (b_after - b_before).abs().max() > 1.0
    model.train()
    with pytest.raises(NotImplementedError):
        model({"image1": img1, "image2": img2})


def test_cuda_neck_matches_the_pytorch_neck_modules():
    """feature_extraction with the CUDA neck (oetr_neck_forward, the default) against the reference's PyTorch modules on the
    SAME weights and backbone features (fp32, TF32 off): features within the neck's fp16-operand bar, boxes within 1e-3."""
    model = _model()
    assert model.neck_mode == "cuda"
    g = torch.Generator().manual_seed(5)
    for shape1, shape2 in (((2, 640, 640, 3), (2, 640, 640, 3)), ((1, 480, 640, 3), (1, 608, 416, 3))):
        img1, img2 = torch.rand(shape1, generator=g).cuda(), torch.rand(shape2, generator=g).cuda()
        with torch.no_grad():
            f1, f2 = model.feature_extraction(img1, img2)[:2]
            b1, b2 = model.forward_dummy(img1, img2)
            model.neck_mode = "torch"
            t1, t2 = model.feature_extraction(img1, img2)[:2]
            c1, c2 = model.forward_dummy(img1, img2)
            model.neck_mode = "cuda"
        for f, t in ((f1, t1), (f2, t2)):
            assert f.shape == t.shape
            std = float(t.std())
            assert float((f - t).abs().max()) < 4e-3 * std and float((f - t).pow(2).mean().sqrt()) < 1e-3 * std
        assert float((b1 - c1).abs().max()) / max(shape1[1:3]) < 1e-3 and float((b2 - c2).abs().max()) / max(shape2[1:3]) < 1e-3
    # the neck handle snapshots its weights and follows in-place updates
    with torch.no_grad():
        model.input_proj2.bias.add_(0.25)
        f1b = model.feature_extraction(img1, img2)[0]
    assert float((f1b - f1).abs().max()) > 0.2


def test_backbone_execution_modes():
    """channels_last / TF32 / bf16 / CUDA-graph execution of the PyTorch trunk against the reference's eager
    fp32 execution; graph replay equals the uncaptured run of the same mode bit for bit and follows new inputs."""
    model = _model()
    g = torch.Generator().manual_seed(7)
    img, img_b = torch.rand((2, 320, 416, 3), generator=g).cuda(), torch.rand((2, 320, 416, 3), generator=g).cuda()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    with torch.no_grad():
        ref = model.backbone(img)
        for mode, tol in (("channels_last", 1e-4), ("tf32", 2e-2), ("bf16", 1e-1)):
            model.backbone.set_execution_mode(mode)
            f = model.backbone(img)
            assert f.shape == ref.shape and f.dtype == torch.float32 and f.is_contiguous() and rel(f, ref) < tol, (mode, rel(f, ref))
            model.backbone.set_execution_mode(mode, graphs=True)
            f1 = model.backbone(img)            # captures
            f2 = model.backbone(img_b)          # replays on new data
            f3 = model.backbone(img)
            assert torch.equal(f1, f3) and rel(f1, ref) < tol and not torch.equal(f1, f2)
            model.backbone.set_execution_mode(mode)
            assert torch.equal(model.backbone(img_b), f2) or rel(model.backbone(img_b), f2) < 1e-5
        model.backbone.set_execution_mode("eager")
        assert torch.equal(model.backbone(img), ref)
    with pytest.raises(ValueError):
        model.backbone.set_execution_mode("int4")
