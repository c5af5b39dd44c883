"""compute-sanitizer driver: a few small forwards of both precisions, flat and per-image tilings, sub-batches and
host requests.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oetr_b200
from oetr_b200 import weights

W = weights.synthetic_hot_path_weights(0)
for precision in ("fp16", "fp32"):
    hot = oetr_b200.OverlapHotPath(W, precision=precision)
    hot.set_chunk_pairs(2)
    for b, fm1, fm2 in ((5, (20, 20), (14, 17)), (2, (7, 13), (11, 5)), (1, (1, 1), (3, 2))):
        f1 = weights.synthetic_features(b, *fm1, seed=3, tag="s1")
        f2 = weights.synthetic_features(b, *fm2, seed=3, tag="s2")
        hw1, hw2 = (fm1[0] * 32, fm1[1] * 32), (fm2[0] * 32, fm2[1] * 32)
        a1, a2 = hot.forward(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), hw1, hw2)
        for _ in range(3):
            h1, h2 = hot.forward_host(f1, f2, hw1, hw2)
        torch.cuda.synchronize()
        print(precision, b, fm1, fm2, float(np.abs(h1 - a1.cpu().numpy()).max()), flush=True)
    hot.poll_error()
    hot.close()
print("done")
