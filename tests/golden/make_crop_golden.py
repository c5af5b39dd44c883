"""Golden vectors of the post-box plumbing from the UNMODIFIED reference function
dloc/core/utils/utils.py:510-564 `tensor_overlap_crop` (which calls cv2.resize INTER_CUBIC) -- needs /root/reference and
cv2 (build container only).   python tests/golden/make_crop_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from crop_cases import CROP_CASES, synthetic_image  # noqa: E402


def main():
    ref_loader.install_dloc()
    import cv2
    from dloc.core.utils.utils import tensor_overlap_crop
    for name, (c, hw1, hw2, box1, box2, extractor, div, seed) in CROP_CASES.items():
        im1, im2 = synthetic_image(c, *hw1, seed), synthetic_image(c, *hw2, seed + 100)
        left, right, r1, r2 = tensor_overlap_crop(torch.from_numpy(im1), torch.tensor([box1]), torch.from_numpy(im2),
                                                  torch.tensor([box2]), extractor, div)
        np.savez_compressed(os.path.join(HERE, "crop_%s.npz" % name), left=left.numpy(), right=right.numpy(),
                            ratio1=np.asarray(r1, np.float64), ratio2=np.asarray(r2, np.float64), cv2_version=cv2.__version__)
        print(name, tuple(left.shape), tuple(right.shape), r1, r2)


if __name__ == "__main__":
    main()
