"""Neck golden cases shared by make_neck_golden.py (needs /root/reference) and the tests.
name: (images, backbone h, backbone w, weight seed, feature seed, weight gain)"""
NECK_CASES = {
    "40x40": (2, 40, 40, 0, 3, 1.0),          # 640 x 640 images
    "30x38": (2, 30, 38, 0, 5, 1.0),          # ragged, even
    "39x37": (1, 39, 37, 1, 7, 1.0),          # odd sizes: the last row / column of the odd phase planes is missing
    "52x52": (1, 52, 52, 0, 9, 1.0),          # 840 x 840
    "8x12": (3, 8, 12, 2, 11, 2.0),           # tiny maps, several images per conv tile, larger weights
}
