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
