"""CPU suite of the neck: the numpy oracle against the committed outputs of the unmodified reference
modules (tests/golden/neck_*.npz, made by make_neck_golden.py), the packed-weight inventory, and the host-only tiling
logic of the C ABI (no compute calls without a GPU)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from neck_cases import NECK_CASES
from oetr_b200 import cabi, weights
from oracle import neck_oracle as nk


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", "neck_%s.npz" % name))


@pytest.mark.parametrize("name", ["8x12", "39x37", "30x38"])
def test_oracle_matches_reference_outputs(name):
    n, h, w, wseed, fseed, gain = NECK_CASES[name]
    W = weights.synthetic_neck_weights(wseed, gain=gain)
    x = weights.synthetic_backbone_features(n, h, w, seed=fseed)
    f = nk.neck(W, x)
    g = _golden(name)
    assert f.shape == (n, 256, h // 2, w // 2)
    # fp64 oracle vs the reference's fp32 arithmetic (K up to 65 536 per output): fp32 summation noise only
    assert np.abs(f[:, ::4] - g["feat_c4"]).max() < 2e-5 * max(1.0, float(g["std"]))
    assert np.allclose(f.sum(axis=(1, 2, 3)), g["sums"], rtol=0, atol=2e-2)


def test_oracle_stride2_conv_is_the_tap_sum_of_its_definition():
    """conv_stride2 against a direct im2col evaluation on a small odd-sized map (padding and floor output size)."""
    rng = np.random.default_rng(0)
    for k in (4, 8, 16):
        x = rng.standard_normal((1, 3, 9, 7))
        wt = rng.standard_normal((2, 3, k, k))
        b = rng.standard_normal(2)
        got = nk.conv_stride2(x, wt, b)
        p = (k - 2) // 2
        xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
        ho, wo = 9 // 2, 7 // 2
        want = np.zeros((1, 2, ho, wo))
        for oy in range(ho):
            for ox in range(wo):
                want[0, :, oy, ox] = np.tensordot(wt, xp[0, :, 2 * oy:2 * oy + k, 2 * ox:2 * ox + k], axes=3) + b
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-10


def test_packed_neck_weights():
    W = weights.synthetic_neck_weights(0)
    packed = weights.pack_neck_weights(W)
    lib = cabi.load_library()
    assert packed.size == weights.NECK_PACKED_COUNT == lib.oetr_neck_packed_weight_count() == 11929088
    with pytest.raises(KeyError):
        weights.pack_neck_weights({k: v for k, v in W.items() if k != "input_proj2.bias"})
    bad = dict(W)
    bad["input_proj.bias"] = np.zeros(3, np.float32)
    with pytest.raises(ValueError):
        weights.pack_neck_weights(bad)


def test_neck_conv_tiling_host_logic():
    """oetr_neck_geometry (host only): box rows <= 128, every output position covered, split-K part counts of the two item
    types (k16 on tile pairs: out[4] // 100; k8 + k4 per tile: out[4] % 100) in range."""
    lib = cabi.load_library()
    out = (ctypes.c_int * 5)()
    expect = {(64, 40, 40): (220, 120, 2, 3, 201), (2, 40, 40): (7, 120, 3, 2, 1604), (32, 52, 52): (208, 104, 2, 2, 101)}
    for n in (1, 2, 3, 7, 32, 64):
        for h, w in ((40, 40), (52, 52), (30, 38), (39, 37), (8, 12), (2, 2), (200, 200), (3, 120)):
            assert lib.oetr_neck_geometry(n, h, w, 148, out) == 0, lib.oetr_neck_last_error()
            tiles, rows, ny, nb, parts = list(out)
            ho, wo = h // 2, w // 2
            assert rows == wo * ny * nb <= 128 and parts // 100 in (1, 2, 4, 8, 16) and parts % 100 in (1, 2, 4, 8)
            assert tiles == -(-ho // ny) * -(-n // nb)
            assert ny * -(-ho // ny) >= ho and nb * -(-n // nb) >= n
            if (n, h, w) in expect:
                assert (tiles, rows, ny, nb, parts) == expect[(n, h, w)], ((n, h, w), tuple(out))
    assert lib.oetr_neck_geometry(1, 1, 40, 148, out) == cabi.OETR_E_SHAPE
    assert lib.oetr_neck_geometry(1, 40, 201, 148, out) == cabi.OETR_E_SHAPE
    assert lib.oetr_neck_geometry(0, 40, 40, 148, out) == cabi.OETR_E_SHAPE


def test_neck_argument_errors_do_not_need_a_gpu():
    lib = cabi.load_library()
    h = ctypes.c_void_p()
    assert lib.oetr_neck_create(None, 0, ctypes.byref(h)) == cabi.OETR_E_ARG
    assert b"null" in lib.oetr_neck_last_error()
    buf = (ctypes.c_float * 4)()
    assert lib.oetr_neck_create(ctypes.cast(buf, ctypes.c_void_p), 4, ctypes.byref(h)) == cabi.OETR_E_ARG
    assert b"expected" in lib.oetr_neck_last_error()
    assert lib.oetr_neck_destroy(None) == 0
    assert lib.oetr_neck_forward(None, None, 1, 40, 40, None, None, 0, None) == cabi.OETR_E_ARG
