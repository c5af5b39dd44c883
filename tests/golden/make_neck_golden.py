"""Golden vectors of the neck from the UNMODIFIED reference modules -- needs /root/reference (build
container only).  Runs input_proj -> patchmerging -> input_proj2 of the reference's OETR (src/model.py:118-124) with the
deterministic synthetic weights / backbone features of oetr_b200.weights and stores every 4th output channel plus
per-image sums.   python tests/golden/make_neck_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from neck_cases import NECK_CASES  # noqa: E402
from oetr_b200 import weights  # noqa: E402


def main():
    ref_loader.install()
    from src.config.default import get_cfg_defaults
    from src.model import build_detectors
    cfg = get_cfg_defaults()
    cfg.OETR.BACKBONE.STRIDE = 32
    torch.manual_seed(0)
    net = build_detectors(cfg.OETR).eval()
    for name, (n, h, w, wseed, fseed, gain) in NECK_CASES.items():
        W = weights.synthetic_neck_weights(wseed, gain=gain)
        sd = net.state_dict()
        for k, v in W.items():
            assert sd[k].shape == v.shape, k
            sd[k] = torch.from_numpy(v)
        net.load_state_dict(sd)
        x = weights.synthetic_backbone_features(n, h, w, seed=fseed)
        with torch.no_grad():
            f = net.input_proj2(net.patchmerging(net.input_proj(torch.from_numpy(x)))).numpy()
        np.savez_compressed(os.path.join(HERE, "neck_%s.npz" % name), feat_c4=f[:, ::4].astype(np.float32),
                            sums=f.astype(np.float64).sum(axis=(1, 2, 3)), abs_sums=np.abs(f.astype(np.float64)).sum(axis=(1, 2, 3)),
                            std=np.float64(f.std()))
        print(name, f.shape, "std %.4f" % f.std())


if __name__ == "__main__":
    main()
