"""CPU suite of SuperGlue's operators: the numpy oracle against the committed outputs of the unmodified
reference module (tests/golden/superglue_*.npz: operator-level cases and the full forward on a sample pair with the in-tree
outdoor weights, which are read from /root/reference when present), and the C ABI's argument checks."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oetr_b200 import cabi
from oracle import superglue_oracle as so

OPS = np.load(os.path.join(ROOT, "tests", "golden", "superglue_ops.npz"))
WEIGHTS = "/root/reference/third_party/SuperGluePretrainedNetwork/models/weights/superglue_outdoor.pth"


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_attention_matches_reference(name):
    q, k, v, want = (OPS["att_%s_%s" % (name, t)] for t in ("q", "k", "v", "out"))
    for b in range(q.shape[0]):
        got = so.attention(q[b].astype(np.float64), k[b].astype(np.float64), v[b].astype(np.float64))
        assert np.abs(got - want[b]).max() < 2e-5


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_optimal_transport_matches_reference(name):
    s, want, it = OPS["ot_%s_in" % name], OPS["ot_%s_out" % name], int(OPS["ot_%s_iters" % name])
    for b in range(s.shape[0]):
        got = so.log_optimal_transport(s[b].astype(np.float64), 2.3, it)
        assert got.shape == want[b].shape and np.abs(got - want[b]).max() < 5e-5


@pytest.mark.skipif(not os.path.exists(WEIGHTS), reason="SuperGlue weights live in the reference tree (build container only)")
def test_full_forward_matches_reference():
    import torch
    g = np.load(os.path.join(ROOT, "tests", "golden", "superglue_pair1.npz"))
    W = {k: v.numpy() for k, v in torch.load(WEIGHTS, map_location="cpu").items()}
    out = so.superglue(W, g["keypoints0"][0], g["keypoints1"][0], g["scores0"][0], g["scores1"][0], g["descriptors0"][0],
                       g["descriptors1"][0], tuple(g["shape0"]), tuple(g["shape1"]), iters=50)
    assert np.abs(out["desc0"] - g["desc0_out"][0]).max() < 2e-4 * np.abs(g["desc0_out"]).max()
    assert np.abs(out["pre_transport"] - g["pre_transport"][0]).max() < 2e-3          # the reference is fp32: 18 layers of noise on |desc| ~ 24
    assert np.abs(out["scores"] - g["scores"][0]).max() < 2e-3
    assert np.array_equal(out["matches0"], g["matches0"][0]) and np.array_equal(out["matches1"], g["matches1"][0])
    assert np.abs(out["matching_scores0"] - g["matching_scores0"][0]).max() < 1e-3


def test_superglue_abi_argument_errors_do_not_need_a_gpu():
    lib = cabi.load_library()
    assert lib.oetr_sg_attention(None, None, None, None, 1, 4, 4, 0, None) == cabi.OETR_E_ARG
    assert lib.oetr_sg_optimal_transport(None, 1.0, 10, None, 1, 4, 4, None, 0, None) == cabi.OETR_E_ARG
    assert lib.oetr_sg_transport_workspace_bytes(2, 10, 20) == 256 + 2 * (32 + 200) * 4
    import torch
    from oetr_b200 import superglue as sg
    model = sg.SuperGlue()
    assert {"kenc.encoder.0.weight", "gnn.layers.17.attn.proj.2.bias", "gnn.layers.0.mlp.1.running_var", "final_proj.weight",
            "bin_score"} <= set(model.state_dict())
    if os.path.exists(WEIGHTS):
        model.load_state_dict(torch.load(WEIGHTS, map_location="cpu"), strict=True)
    if not torch.cuda.is_available():
        with pytest.raises(cabi.OetrError):
            sg.attention(torch.zeros(1, 64, 4, 5), torch.zeros(1, 64, 4, 5), torch.zeros(1, 64, 4, 5))
        with pytest.raises(cabi.OetrError):
            sg.log_optimal_transport(torch.zeros(1, 5, 5), torch.tensor(1.0), 3)
