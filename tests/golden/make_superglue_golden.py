"""Golden vectors of SuperGlue from the UNMODIFIED reference module with its in-tree outdoor weights -- needs
/root/reference (build container only).  Inputs: SuperPoint (reference code + in-tree weights) on the first sample pair of
tools/pipeline_sample.py, capped at 96 keypoints to keep the fixture small; plus operator-level cases (attention,
log_optimal_transport) on seeded random inputs.   python tests/golden/make_superglue_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, "/root/reference/third_party/SuperGluePretrainedNetwork")


def main():
    import pipeline_sample as ps
    from models import superglue as ref
    from models.superpoint import SuperPoint
    sp = SuperPoint({"nms_radius": 4, "keypoint_threshold": 0.005, "max_keypoints": 96}).eval()
    sg = ref.SuperGlue({"weights": "outdoor", "sinkhorn_iterations": 50, "match_threshold": 0.2}).eval()
    (g0, _, _), (g1, _, _) = ps.read_pair_images("pair_1")
    with torch.no_grad():
        p0, p1 = sp({"image": g0}), sp({"image": g1})
        data = {"image0": g0, "image1": g1}
        data.update({k + "0": torch.stack(v) for k, v in p0.items()})
        data.update({k + "1": torch.stack(v) for k, v in p1.items()})
        out = sg(data)
        # the full assignment matrix, recomputed with the reference's own functions
        k0 = ref.normalize_keypoints(data["keypoints0"], g0.shape)
        k1 = ref.normalize_keypoints(data["keypoints1"], g1.shape)
        d0 = data["descriptors0"] + sg.kenc(k0, data["scores0"])
        d1 = data["descriptors1"] + sg.kenc(k1, data["scores1"])
        d0, d1 = sg.gnn(d0, d1)
        sc = torch.einsum("bdn,bdm->bnm", sg.final_proj(d0), sg.final_proj(d1)) / 256 ** 0.5
        Z = ref.log_optimal_transport(sc, sg.bin_score, iters=50)
    np.savez_compressed(os.path.join(HERE, "superglue_pair1.npz"), shape0=np.asarray(g0.shape), shape1=np.asarray(g1.shape),
                        **{k: v.numpy() for k, v in data.items() if not k.startswith("image")},
                        matches0=out["matches0"].numpy(), matches1=out["matches1"].numpy(),
                        matching_scores0=out["matching_scores0"].numpy(), pre_transport=sc.numpy(), scores=Z.numpy(),
                        desc0_out=d0.numpy())
    print("pair1", data["keypoints0"].shape, data["keypoints1"].shape, int((out["matches0"] > -1).sum()), "matches")
    g = torch.Generator().manual_seed(0)
    ops = {}
    for name, (b, n, m) in {"a": (2, 40, 90), "b": (1, 64, 64), "c": (1, 1, 70), "d": (1, 130, 3)}.items():
        q, k, v = torch.randn(b, 64, 4, n, generator=g) * 1.5, torch.randn(b, 64, 4, m, generator=g) * 1.5, torch.randn(b, 64, 4, m, generator=g)
        ops["att_%s_q" % name], ops["att_%s_k" % name], ops["att_%s_v" % name] = q.numpy(), k.numpy(), v.numpy()
        ops["att_%s_out" % name] = ref.attention(q, k, v)[0].numpy()
    for name, (m, n, it) in {"a": (50, 77, 20), "b": (1, 9, 5), "c": (130, 40, 100)}.items():
        s = torch.randn(2, m, n, generator=g) * 3.0
        ops["ot_%s_in" % name] = s.numpy()
        ops["ot_%s_out" % name] = ref.log_optimal_transport(s, torch.tensor(2.3), it).numpy()
        ops["ot_%s_iters" % name] = np.asarray(it)
    np.savez_compressed(os.path.join(HERE, "superglue_ops.npz"), **ops)
    print("ops", len(ops))


if __name__ == "__main__":
    main()
