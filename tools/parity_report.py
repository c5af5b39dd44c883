"""Parity report of the CUDA hot path against the committed outputs of the real reference (tests/golden/*.npz, fp64 run):
max |box - box_ref| / image side (unclamped boxes) for every golden case, both precision paths.  GPU only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import oetr_b200  # noqa: E402
from cases import CASES, MASK_CASES, STRESS_CASES  # noqa: E402
from conftest import load_case, load_masks, load_stress_case  # noqa: E402


def run(W, f1, f2, hw1, hw2, attention, precision, m1=None, m2=None):
    hot = oetr_b200.OverlapHotPath(W, attention=attention, precision=precision)
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()
    b1, b2 = hot.forward(t(f1), t(f2), hw1, hw2, clamp=False, mask1=t(m1), mask2=t(m2))
    torch.cuda.synchronize()
    out = b1.cpu().numpy(), b2.cpu().numpy()
    hot.close()
    return out


print("%-22s %-7s %12s %12s   (box error / image side vs the reference run in fp64; bar 1e-3)" % ("case", "attn", "fp32 path", "fp16 path"))
for name in sorted(CASES) + sorted(MASK_CASES) + sorted(STRESS_CASES):
    if name in STRESS_CASES:
        W, f1, f2, c, g = load_stress_case(name)
        hw1, hw2, attention, m1, m2 = c["hw1"], c["hw2"], "linear", None, None
    else:
        W, f1, f2, (b, fm1, fm2, hw1, hw2, attention, _, _), g = load_case(name)
        m1, m2 = load_masks(name) if name in MASK_CASES else (None, None)
    errs = []
    for precision in ("fp32", "fp16"):
        b1, b2 = run(W, f1, f2, hw1, hw2, attention, precision, m1, m2)
        errs.append(max(np.abs(b1 - g["box1_raw_f64"]).max() / max(hw1), np.abs(b2 - g["box2_raw_f64"]).max() / max(hw2)))
    print("%-22s %-7s %12.2e %12.2e" % (name, attention, errs[0], errs[1]))
