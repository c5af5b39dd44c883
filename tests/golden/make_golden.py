"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (read-only /root/reference) on
deterministic synthetic weights and features.  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are NOT stored: they are regenerated bit-exactly from (name-keyed Philox) seeds by
imagematching-oetr_b200/weights.py; the files hold only the reference's outputs at every stage boundary
(fp32 = the reference as shipped, fp64 = the same modules in double precision, used as the noise floor).
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from cases import CASES, MASK_CASES, MEMORY_STRIDE, STRESS_CASES  # noqa: E402

from oetr_b200 import weights  # noqa: E402


def load_synthetic(net, seed, **variant):
    sd = weights.synthetic_hot_path_weights(seed, include_unused=True, **variant)
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("backbone.", "input_proj", "patchmerging.")) for k in missing), missing
    return sd


def run_reference(net, full_tf, feat1, feat2, hw1, hw2, attention, dtype, mask1=None, mask2=None):
    """reference src/model.py:240-250 on precomputed features"""
    if mask1 is not None:
        mask1, mask2 = torch.from_numpy(mask1).to(dtype), torch.from_numpy(mask2).to(dtype)
    from src.models.utils import box_tlbr_to_xyxy
    f1 = torch.from_numpy(feat1).to(dtype)
    f2 = torch.from_numpy(feat2).to(dtype)
    net.h1, net.w1 = hw1
    net.h2, net.w2 = hw2
    hf1, wf1 = f1.shape[2:]
    hf2, wf2 = f2.shape[2:]
    with torch.no_grad():
        pos1, pos2 = net.pos_encoding(f1), net.pos_encoding(f2)
        if attention == "linear":
            hs1, hs2, m1, m2 = net.feature_correlation(f1, f2, pos1, pos2, mask1, mask2)
        else:
            hs1, hs2, m1, m2 = full_tf(f1, f2, net.query_embed1.weight, net.query_embed2.weight, pos1, pos2,
                                       None, None)
        cxy1, cxy2 = net.center_estimation(hs1, hs2, m1, m2, hf1, wf1, hf2, wf2, mask1, mask2)
        tlbr1, tlbr2 = net.size_regression(hs1, hs2)
        box1 = box_tlbr_to_xyxy(cxy1, tlbr1, max_h=hw1[0], max_w=hw1[1])
        box2 = box_tlbr_to_xyxy(cxy2, tlbr2, max_h=hw2[0], max_w=hw2[1])
        raw1, raw2, _, _ = net.obtain_overlap_bbox(cxy1, tlbr1, cxy2, tlbr2)
    out = dict(hs1=hs1[:, 0], hs2=hs2[:, 0], cxy1=cxy1, cxy2=cxy2, tlbr1=tlbr1, tlbr2=tlbr2, box1=box1,
               box2=box2, box1_raw=raw1, box2_raw=raw2,
               memory1_sub=m1[:, ::MEMORY_STRIDE], memory2_sub=m2[:, ::MEMORY_STRIDE],
               memory1_sum=m1.double().sum(dim=(1, 2)), memory2_sum=m2.double().sum(dim=(1, 2)),
               memory1_abs=m1.double().abs().sum(dim=(1, 2)), memory2_abs=m2.double().abs().sum(dim=(1, 2)))
    return {k: v.cpu().numpy() for k, v in out.items()}


def make_stress(net):
    for name, c in STRESS_CASES.items():
        load_synthetic(net, c["wseed"], ln_gain=c["ln_gain"], head_default_init=c["head_default_init"])
        feat1 = weights.synthetic_features(c["batch"], *c["fm1"], seed=c["fseed"], tag="feat1", scale=c["feat_scale"])
        feat2 = weights.synthetic_features(c["batch"], *c["fm2"], seed=c["fseed"], tag="feat2", scale=c["feat_scale"])
        net.float()
        o32 = run_reference(net, None, feat1, feat2, c["hw1"], c["hw2"], "linear", torch.float32)
        net.double()
        o64 = run_reference(net, None, feat1, feat2, c["hw1"], c["hw2"], "linear", torch.float64)
        net.float()
        blob = {k: v for k, v in o32.items()}
        blob.update({k + "_f64": v for k, v in o64.items() if not k.startswith("memory")})
        if c["batch"] > 4:                                # big batch: keep memory of 4 pairs only
            for k in ("memory1_sub", "memory2_sub"):
                blob[k] = blob[k][::8]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print("%-20s box1_raw=%s  |fp32-fp64|=%.2e  tlbr1=%s" % (name, o64["box1_raw"][0].round(3),
              np.abs(o32["box1_raw"] - o64["box1_raw"]).max(), o64["tlbr1"][0].round(4)))


def main():
    net = ref_loader.build_reference_oetr(seed=0)
    if "--stress-only" in sys.argv:                    # leaves the other committed files untouched
        make_stress(net)
        return
    only_masked = "--masked-only" in sys.argv          # leaves the other committed files untouched
    if not only_masked:
        np.savez_compressed(os.path.join(HERE, "pe_table.npz"),
                            pe_sub=net.pos_encoding.pe[0, :, ::9, ::7].numpy(),
                            pe_corner=net.pos_encoding.pe[0, :, :3, :3].numpy())
    for name, (b, fm1, fm2, hw1, hw2, attention, wseed, fseed) in MASK_CASES.items():
        load_synthetic(net, wseed)
        feat1 = weights.synthetic_features(b, *fm1, seed=fseed, tag="feat1")
        feat2 = weights.synthetic_features(b, *fm2, seed=fseed, tag="feat2")
        mk1, mk2 = weights.synthetic_mask(b, *fm1, tag="mask1"), weights.synthetic_mask(b, *fm2, tag="mask2")
        net.float()
        o32 = run_reference(net, None, feat1, feat2, hw1, hw2, attention, torch.float32, mk1, mk2)
        net.double()
        o64 = run_reference(net, None, feat1, feat2, hw1, hw2, attention, torch.float64, mk1, mk2)
        net.float()
        blob = {k: v for k, v in o32.items()}
        blob.update({k + "_f64": v for k, v in o64.items() if not k.startswith("memory")})
        blob["memory1_sub_f64"] = o64["memory1_sub"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print("%-16s box1_raw=%s  |fp32-fp64|=%.2e" % (name, o64["box1_raw"][0].round(3),
                                                       np.abs(o32["box1_raw"] - o64["box1_raw"]).max()))
    if only_masked:
        return
    for name, (b, fm1, fm2, hw1, hw2, attention, wseed, fseed) in CASES.items():
        load_synthetic(net, wseed)
        full_tf = ref_loader.build_reference_full_transformer(net) if attention == "full" else None
        feat1 = weights.synthetic_features(b, *fm1, seed=fseed, tag="feat1")
        feat2 = weights.synthetic_features(b, *fm2, seed=fseed, tag="feat2")
        net.float()
        o32 = run_reference(net, full_tf, feat1, feat2, hw1, hw2, attention, torch.float32)
        net.double()
        if full_tf is not None:
            full_tf.double()
        o64 = run_reference(net, full_tf, feat1, feat2, hw1, hw2, attention, torch.float64)
        net.float()
        blob = {k: v for k, v in o32.items()}
        blob.update({k + "_f64": v for k, v in o64.items() if not k.startswith("memory")})
        blob["memory1_sub_f64"] = o64["memory1_sub"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        d = np.abs(o32["box1_raw"] - o64["box1_raw"]).max()
        print("%-16s box1_raw=%s  |fp32-fp64|=%.2e  tlbr1=%s" % (name, o64["box1_raw"][0].round(3), d,
                                                                 o64["tlbr1"][0].round(4)))


    make_stress(net)


if __name__ == "__main__":
    main()
