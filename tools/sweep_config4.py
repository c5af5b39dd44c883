"""BASELINE config 4 (840x840 pairs, batch 16: 26x26 maps, 676 tokens) and config 2 (640x640, batch 32): device
time of the hot path per encoder kernel variant and sub-batch size.  Usage: sweep_config4.py [840|640]
   OETR_ENC=1|2 selects k_enc (one CTA per 128-token tile) / k_enc2 (CTA pairs, tcgen05 cta_group::2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oetr_b200
from oetr_b200 import weights

side = int(sys.argv[1]) if len(sys.argv) > 1 else 840
b, fm = (16, 26) if side == 840 else (32, 20)
W = weights.synthetic_hot_path_weights(0)
f1 = torch.from_numpy(weights.synthetic_features(b, fm, fm, seed=4, tag="l1")).cuda()
f2 = torch.from_numpy(weights.synthetic_features(b, fm, fm, seed=4, tag="l2")).cuda()
hot = oetr_b200.OverlapHotPath(W, precision="fp16")
for cp in (0, 4, 8):
    hot.set_chunk_pairs(cp)
    for _ in range(3):
        hot.forward(f1, f2, (side, side), (side, side))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        hot.forward(f1, f2, (side, side), (side, side))
    e1.record()
    torch.cuda.synchronize()
    hot.poll_error()
    ms = e0.elapsed_time(e1) / 20
    print("%dx%d batch %d  OETR_ENC=%s  sub-batch pairs=%d: %.3f ms/step  %.0f pairs/s" % (
        side, side, b, os.environ.get("OETR_ENC", "1"), cp, ms, b / ms * 1e3), flush=True)
