"""Golden-vector case list shared by make_golden.py (generator, needs /root/reference) and the tests."""

# name: (batch, (hf1, wf1), (hf2, wf2), (img_h1, img_w1), (img_h2, img_w2), attention, weight_seed, feat_seed)
CASES = {
    "b2_640": (2, (20, 20), (20, 20), (640, 640), (640, 640), "linear", 0, 1),
    "ragged_640x480": (1, (20, 20), (15, 20), (640, 640), (480, 640), "linear", 0, 2),
    "b1_840": (1, (26, 26), (26, 26), (840, 840), (840, 840), "linear", 0, 3),
    "tiny_b3": (3, (5, 7), (4, 6), (160, 224), (128, 192), "linear", 3, 4),
    "stride31_600": (1, (19, 19), (19, 19), (600, 600), (600, 600), "linear", 0, 5),
    "full_640": (1, (20, 20), (20, 20), (640, 640), (640, 640), "full", 0, 6),
    "full_ragged": (1, (20, 20), (15, 20), (640, 640), (480, 640), "full", 0, 7),
}
MEMORY_STRIDE = 7      # golden files keep every 7th memory token (all channels) plus whole-tensor sums

# the same tuple; both images carry a padding-style float mask from weights.synthetic_mask (masks are
# reachable only by direct callers of forward_dummy / feature_correlation)
MASK_CASES = {
    "masked_640x480": (2, (20, 20), (15, 20), (640, 640), (480, 640), "linear", 0, 8),
    "masked_tiny": (3, (5, 7), (4, 6), (160, 224), (128, 192), "linear", 3, 9),
}

# Stress cases: trained-like scales.  name -> dict(batch, fm1, fm2, hw1, hw2, wseed, fseed,
# feat_scale, ln_gain (lo, hi) or None, head_default_init).  All linear attention.  b32_640 is BASELINE config 2's full
# batch; its golden file keeps every box and a sub-sampled memory.
STRESS_CASES = {
    "stress_feat_x3": dict(batch=2, fm1=(20, 20), fm2=(20, 20), hw1=(640, 640), hw2=(640, 640), wseed=0, fseed=11,
                           feat_scale=3.0, ln_gain=None, head_default_init=False),
    "stress_feat_x0.1": dict(batch=2, fm1=(20, 20), fm2=(15, 20), hw1=(640, 640), hw2=(480, 640), wseed=0, fseed=12,
                             feat_scale=0.1, ln_gain=None, head_default_init=False),
    "stress_ln_gain3": dict(batch=2, fm1=(20, 20), fm2=(20, 20), hw1=(640, 640), hw2=(640, 640), wseed=0, fseed=13,
                            feat_scale=1.0, ln_gain=(0.5, 3.0), head_default_init=False),
    "stress_default_head": dict(batch=2, fm1=(20, 20), fm2=(20, 20), hw1=(640, 640), hw2=(640, 640), wseed=0, fseed=14,
                                feat_scale=1.0, ln_gain=(0.7, 2.0), head_default_init=True),
    "b32_640": dict(batch=32, fm1=(20, 20), fm2=(20, 20), hw1=(640, 640), hw2=(640, 640), wseed=0, fseed=15,
                    feat_scale=1.0, ln_gain=None, head_default_init=False),
}
