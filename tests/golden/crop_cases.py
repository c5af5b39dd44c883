"""Crop / resize golden cases shared by make_crop_golden.py (needs /root/reference + cv2) and the tests.
name: (channels, (H1, W1), (H2, W2), box1 xyxy float, box2 xyxy float, extractor_name, size_divisor, seed)"""
CROP_CASES = {
    "gray_superpoint": (1, (120, 160), (120, 160), (21.7, 10.2, 131.9, 97.5), (40.1, 33.3, 155.8, 119.6), "superpoint", 1, 1),
    "gray_div8": (1, (150, 111), (96, 128), (5.0, 7.9, 100.2, 140.0), (0.0, 0.0, 128.0, 96.0), "superpoint", 8, 2),
    "rgb_disk": (3, (96, 128), (128, 96), (10.5, 3.5, 90.5, 64.0), (3.0, 30.0, 77.7, 120.4), "disk", 1, 3),
    "tall_box": (1, (128, 128), (128, 128), (60.0, 2.0, 75.9, 126.0), (1.0, 50.0, 127.0, 70.5), "superpoint", 1, 4),
    "box_past_edge": (1, (100, 140), (100, 140), (100.0, 60.0, 150.0, 120.0), (0.0, 0.0, 140.0, 100.0), "superpoint", 1, 5),
}


def synthetic_image(channels, h, w, seed):
    """Smooth + textured deterministic image [1,C,h,w] float32 in [0,1] (numpy only)."""
    import numpy as np
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    out = []
    for c in range(channels):
        a = 0.5 + 0.25 * np.sin(x / (7.0 + c) + 0.3 * seed) * np.cos(y / (5.0 + 2 * c)) + 0.2 * (rng.integers(0, 1 << 16, (h, w)) / 65536.0 - 0.5)
        out.append(np.clip(a, 0.0, 1.0))
    return np.stack(out)[None].astype(np.float32)
