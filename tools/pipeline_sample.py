"""BASELINE config 5 substitute: the reference's overlap -> crop -> SuperPoint -> SuperGlue pipeline
(evaluation.py:77-140) on the three sample pairs that ship inside the reference tree
(third_party/D2Net/qualitative/images/pair_{1,2,3}), reference against replacement.  Pose AUC on MegaDepth / IMC is NOT
reproducible here (no datasets, no trained OETR checkpoint in the tree); what this measures is whether the replacement
hands the matcher the same boxes and the same crops: box agreement, crop agreement and match counts.

Both sides use the SAME deterministic weights for the whole OETR model (name-keyed generator below: there is no trained
checkpoint), the in-tree SuperPoint / SuperGlue-outdoor weights, grayscale matching images at their native size and
640 x 640 overlap images (read_overlap_image, dloc/core/utils/utils.py:271-340).

Phases (the reference tree exists only in the build container, the GPU only on the GPU box):
  --phase prepare    copy the six sample JPEGs to oracle/_ref/samples/ (git-ignored, travels to the GPU box)
  --phase reference  build container, CPU: unmodified reference OETR + tensor_overlap_crop (cv2) + SuperPoint + SuperGlue
                     -> tests/golden/pipeline_ref.json
  --phase gpu        GPU box: oetr_b200 plugin model (CUDA neck + hot path) + device crop (oetr_crop_resize)
                     -> gpurun_out/pipeline_gpu.npz (boxes + crops)
  --phase compare    build container: reference SuperPoint + SuperGlue on the replacement's crops, agreement report
                     -> profiles/r02_pipeline_sample.json
"""
import argparse
import json
import math
import os
import shutil
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SAMPLES = os.path.join(ROOT, "oracle", "_ref", "samples")
REF = os.environ.get("OETR_REFERENCE_ROOT", "/root/reference")
PAIRS = ("pair_1", "pair_2", "pair_3")
OVERLAP_SIZE = 640
SP_CONF = {"nms_radius": 4, "keypoint_threshold": 0.005, "max_keypoints": 2048}
SG_CONF = {"weights": "outdoor", "sinkhorn_iterations": 20, "match_threshold": 0.2}


def synthetic_state_dict(model):
    """Deterministic weights for EVERY tensor of an OETR model, keyed by the state_dict name (identical for the reference
    class and the mirror, which share names and shapes): kaiming-scale convolutions / linears, identity BatchNorm statistics
    with a damped last BatchNorm per bottleneck (keeps the random ResNet's activations in range), LayerNorm/GroupNorm at 1 / 0."""
    out = {}
    for name, t in model.state_dict().items():
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            v = torch.zeros_like(t)
        elif name.endswith("running_mean"):
            v = torch.zeros_like(t)
        elif name.endswith("running_var"):
            v = torch.ones_like(t)
        elif t.dim() == 1:
            is_norm_w = name.endswith("weight") and (".bn" in name or "norm" in name or "downsample.1" in name or name.startswith("heatmap_conv.1"))
            if is_norm_w:
                v = torch.full_like(t, 0.5 if ".bn3." in name else 1.0)
            else:
                v = (torch.rand(t.shape, generator=g) - 0.5) * 0.2
        elif "query_embed" in name:
            v = (torch.rand(t.shape, generator=g) - 0.5) * 3.0
        else:
            fan_in = int(np.prod(t.shape[1:]))
            gain = math.sqrt(2.0) if name.startswith("backbone.") else 1.0
            v = (torch.rand(t.shape, generator=g) - 0.5) * (2.0 * gain * math.sqrt(3.0 / fan_in))
        out[name] = v.to(t.dtype)
    return out


def read_pair_images(pair):
    """read_overlap_image (dloc/core/utils/utils.py:271-340) with overlap=True, grayscale=True, align='', resize=[-1],
    overlap resize 640: (gray [1,1,H,W], overlap image [1,640,640,3] RGB, overlap_scales) per image, all in [0,1]."""
    import cv2
    out = []
    for i in (1, 2):
        path = os.path.join(SAMPLES, pair, "%d.jpg" % i)
        image = cv2.imread(path, cv2.IMREAD_COLOR)
        if image is None:
            raise FileNotFoundError(path + " (run --phase prepare in the build container first)")
        image = image[:, :, ::-1].astype(np.float32)
        h, w = image.shape[:2]
        overlap_image = cv2.resize(image, (OVERLAP_SIZE, OVERLAP_SIZE))
        gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
        out.append((torch.from_numpy(gray[None, None] / 255.0).float(), torch.from_numpy(overlap_image[None] / 255.0).float(),
                    (float(w) / OVERLAP_SIZE, float(h) / OVERLAP_SIZE)))
    return out


def load_matchers():
    sys.path.insert(0, os.path.join(REF, "third_party", "SuperGluePretrainedNetwork"))
    from models.superglue import SuperGlue
    from models.superpoint import SuperPoint
    return SuperPoint(SP_CONF).eval(), SuperGlue(SG_CONF).eval()


def match_count(sp, sg, left, right):
    with torch.no_grad():
        p0, p1 = sp({"image": left}), sp({"image": right})
        data = {"image0": left, "image1": right}
        data.update({k + "0": torch.stack(v) for k, v in p0.items()})
        data.update({k + "1": torch.stack(v) for k, v in p1.items()})
        m = sg(data)
    return int(p0["keypoints"][0].shape[0]), int(p1["keypoints"][0].shape[0]), int((m["matches0"][0] > -1).sum())


def scaled_boxes(b0, b1, s0, s1):
    return b0 * torch.tensor(s0 + s0, device=b0.device), b1 * torch.tensor(s1 + s1, device=b1.device)      # evaluation.py:72-84


def phase_prepare():
    for pair in PAIRS:
        os.makedirs(os.path.join(SAMPLES, pair), exist_ok=True)
        for i in (1, 2):
            shutil.copyfile(os.path.join(REF, "third_party", "D2Net", "qualitative", "images", pair, "%d.jpg" % i),
                            os.path.join(SAMPLES, pair, "%d.jpg" % i))
    wdir = os.path.join(ROOT, "oracle", "_ref", "weights")          # in-tree matcher weights for the GPU tests of oetr_b200.superglue
    os.makedirs(wdir, exist_ok=True)
    shutil.copyfile(os.path.join(REF, "third_party", "SuperGluePretrainedNetwork", "models", "weights", "superglue_outdoor.pth"),
                    os.path.join(wdir, "superglue_outdoor.pth"))
    print("samples in", SAMPLES)


def phase_reference(out_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_loader
    ref_loader.install_dloc()
    from dloc.core.utils.utils import tensor_overlap_crop
    from src.config.default import get_cfg_defaults
    from src.model import build_detectors
    cfg = get_cfg_defaults()
    cfg.OETR.BACKBONE.STRIDE = 32
    net = build_detectors(cfg.OETR).eval()
    net.load_state_dict(synthetic_state_dict(net))
    sp, sg = load_matchers()
    res = {}
    for pair in PAIRS:
        (g0, o0, s0), (g1, o1, s1) = read_pair_images(pair)
        with torch.no_grad():
            b0, b1 = net.forward_dummy(o0, o1)
        b0, b1 = scaled_boxes(b0, b1, s0, s1)
        left, right, r0, r1 = tensor_overlap_crop(g0, b0, g1, b1, "superpoint", 1)
        k0, k1, n = match_count(sp, sg, left, right)
        d0, d1, nd = match_count(sp, sg, g0, g1)
        res[pair] = {"bbox0": b0[0].tolist(), "bbox1": b1[0].tolist(), "crop0": list(left.shape[2:]), "crop1": list(right.shape[2:]),
                     "ratio0": r0, "ratio1": r1, "keypoints": [k0, k1], "matches": n, "matches_without_overlap": nd,
                     "crop0_mean": float(left.mean()), "crop1_mean": float(right.mean())}
        print(pair, res[pair])
    json.dump(res, open(out_path, "w"), indent=1)


def phase_gpu(out_path):
    import oetr_b200
    from oetr_b200.dloc.core.utils import utils as U
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = oetr_b200.build_detectors(oetr_b200.get_cfg_defaults().OETR)
    net.load_state_dict(synthetic_state_dict(net))
    net = net.cuda().eval()
    out = {}
    for pair in PAIRS:
        (g0, o0, s0), (g1, o1, s1) = read_pair_images(pair)
        b0, b1 = net.forward_dummy(o0.cuda(), o1.cuda())
        b0, b1 = scaled_boxes(b0, b1, s0, s1)
        assert U.overlap_gate(b0, b1)
        left, right, r0, r1 = U.tensor_overlap_crop(g0.cuda(), b0, g1.cuda(), b1, "superpoint", 1)
        out[pair + "_bbox0"], out[pair + "_bbox1"] = b0.cpu().numpy(), b1.cpu().numpy()
        out[pair + "_left"], out[pair + "_right"] = left.cpu().numpy(), right.cpu().numpy()
        out[pair + "_ratio"] = np.asarray([r0, r1], np.float64)
        print(pair, b0.tolist(), b1.tolist(), tuple(left.shape), tuple(right.shape))
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **out)


def phase_compare(ref_path, gpu_path, out_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_loader
    ref_loader.install_dloc()
    from dloc.core.utils.utils import tensor_overlap_crop
    ref = json.load(open(ref_path))
    gpu = np.load(gpu_path)
    sp, sg = load_matchers()
    rep = {"note": "reference = unmodified OETR + cv2 crop on the CPU; replacement = oetr_b200 on a B200 (CUDA neck + hot path + "
                   "device crop); the SAME deterministic (untrained) OETR weights on both sides, in-tree SuperPoint / SuperGlue-"
                   "outdoor weights; matcher run by the reference's code on the CPU for both.  Pose AUC is not reproducible here.",
           "pairs": {}}
    for pair in PAIRS:
        (g0, _, _), (g1, _, _) = read_pair_images(pair)
        b0, b1 = torch.from_numpy(gpu[pair + "_bbox0"]), torch.from_numpy(gpu[pair + "_bbox1"])
        left, right = torch.from_numpy(gpu[pair + "_left"]), torch.from_numpy(gpu[pair + "_right"])
        k0, k1, n = match_count(sp, sg, left, right)
        # the reference's own crop of the REPLACEMENT's boxes: isolates the crop kernel from box differences
        rl, rr, _, _ = tensor_overlap_crop(g0, b0, g1, b1, "superpoint", 1)
        r = ref[pair]
        side0, side1 = max(g0.shape[2:]), max(g1.shape[2:])
        rep["pairs"][pair] = {
            "box_diff_over_side": [float(np.abs(b0[0].numpy() - np.asarray(r["bbox0"])).max() / side0),
                                   float(np.abs(b1[0].numpy() - np.asarray(r["bbox1"])).max() / side1)],
            "bbox0": b0[0].tolist(), "bbox0_reference": r["bbox0"], "bbox1": b1[0].tolist(), "bbox1_reference": r["bbox1"],
            "crop_shapes": [list(left.shape[2:]), list(right.shape[2:])], "crop_shapes_reference": [r["crop0"], r["crop1"]],
            "crop_max_abs_diff_vs_cv2_on_same_boxes": [float((left - rl).abs().max()) if left.shape == rl.shape else None,
                                                       float((right - rr).abs().max()) if right.shape == rr.shape else None],
            "keypoints": [k0, k1], "keypoints_reference": r["keypoints"], "matches": n, "matches_reference": r["matches"],
            "matches_without_overlap_reference": r["matches_without_overlap"]}
        print(pair, json.dumps(rep["pairs"][pair]))
    json.dump(rep, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", required=True, choices=["prepare", "reference", "gpu", "compare"])
    ap.add_argument("--ref", default=os.path.join(ROOT, "tests", "golden", "pipeline_ref.json"))
    ap.add_argument("--gpu-out", default=os.path.join(ROOT, "gpurun_out", "pipeline_gpu.npz"))
    ap.add_argument("--report", default=os.path.join(ROOT, "profiles", "r02_pipeline_sample.json"))
    a = ap.parse_args()
    if a.phase == "prepare":
        phase_prepare()
    elif a.phase == "reference":
        phase_reference(a.ref)
    elif a.phase == "gpu":
        phase_gpu(a.gpu_out)
    else:
        phase_compare(a.ref, a.gpu_out, a.report)
