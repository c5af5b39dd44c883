"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle and the committed reference outputs.

Tolerances (north star: boxes within 1e-3 relative of the fp32 reference):
  fp32 path : box error / image side < 2e-5   (fp32 summation-order noise; the reference's own fp32-vs-fp64 gap
              is ~2e-6), intermediates < 5e-5 relative
  fp16 path : box error / image side < 1e-3   (the stated bar), memory/hs < 3e-3 relative (max-norm)
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from cases import CASES, MASK_CASES, MEMORY_STRIDE, STRESS_CASES
from conftest import load_case, load_masks, load_stress_case, rel_err
from oracle import oetr_oracle as orc

pytestmark = pytest.mark.gpu

import oetr_b200  # noqa: E402
from oetr_b200 import cabi, weights  # noqa: E402

TOL = {"fp32": dict(box=2e-5, mid=5e-5), "fp16": dict(box=1e-3, mid=3e-3)}


def _run(W, f1, f2, hw1, hw2, attention, precision, clamp):
    hot = oetr_b200.OverlapHotPath(W, attention=attention, precision=precision)
    b1, b2, dbg = hot.forward(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), hw1, hw2, clamp=clamp,
                              debug=True)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in dbg.items()}
    out["box1"], out["box2"] = b1.cpu().numpy(), b2.cpu().numpy()
    out["launches"] = hot.last_launch_count
    hot.close()
    return out


def _cases(precision):
    return sorted(CASES)            # both attention modes run on both precision paths (round 2: tcgen05 full attention)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_stage_parity_with_oracle_and_golden(name, precision):
    W, f1, f2, (b, fm1, fm2, hw1, hw2, attention, _, _), g = load_case(name)
    want = orc.hot_path(W, f1, f2, hw1, hw2, attention=attention)
    got = _run(W, f1, f2, hw1, hw2, attention, precision, clamp=False)
    tol = TOL[precision]
    assert got["launches"] > 0
    for k in ("memory1", "memory2", "hs1", "hs2", "tlbr1", "tlbr2"):
        assert rel_err(got[k], want[k]) < tol["mid"], (k, rel_err(got[k], want[k]))
    for i, hw in ((1, hw1), (2, hw2)):
        side = max(hw)
        assert np.abs(got["cxy%d" % i] - want["cxy%d" % i]).max() / side < tol["box"]
        assert np.abs(got["box%d" % i] - want["box%d_raw" % i]).max() / side < tol["box"]
        # committed outputs of the real reference (fp32 as shipped)
        assert np.abs(got["box%d" % i] - g["box%d_raw" % i]).max() / side < tol["box"] + 2e-5
        assert rel_err(got["memory%d" % i][:, ::MEMORY_STRIDE], g["memory%d_sub" % i]) < tol["mid"] + 2e-5
    clamped = _run(W, f1, f2, hw1, hw2, attention, precision, clamp=True)
    for i, hw in ((1, hw1), (2, hw2)):
        assert np.abs(clamped["box%d" % i] - g["box%d" % i]).max() / max(hw) < tol["box"] + 2e-5
        assert clamped["box%d" % i].min() >= 0 and clamped["box%d" % i][:, 0::2].max() <= hw[1]


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_edge_shapes(precision):
    W = weights.synthetic_hot_path_weights(0)
    # 1x1 maps, a single row, a non-multiple-of-anything map, and the PositionEncodingSine maximum (100x100)
    for b, fm1, fm2 in ((1, (1, 1), (1, 1)), (2, (1, 9), (3, 1)), (1, (7, 13), (11, 5)), (1, (100, 100), (2, 2))):
        hw1, hw2 = (fm1[0] * 32, fm1[1] * 32), (fm2[0] * 32, fm2[1] * 32)
        f1 = weights.synthetic_features(b, *fm1, seed=5, tag="e1")
        f2 = weights.synthetic_features(b, *fm2, seed=5, tag="e2")
        want = orc.hot_path(W, f1, f2, hw1, hw2)
        got = _run(W, f1, f2, hw1, hw2, "linear", precision, clamp=False)
        for i, hw in ((1, hw1), (2, hw2)):
            err = np.abs(got["box%d" % i] - want["box%d_raw" % i]).max() / max(hw)
            assert err < TOL[precision]["box"], (fm1, fm2, i, err)


def test_empty_batch_and_error_codes():
    W = weights.synthetic_hot_path_weights(0)
    hot = oetr_b200.OverlapHotPath(W, precision="fp32")
    e = torch.empty(0, 256, 4, 4, device="cuda")
    b1, b2 = hot.forward(e, e, (128, 128), (128, 128))
    assert b1.shape == (0, 4) and b2.shape == (0, 4)
    with pytest.raises(cabi.OetrError) as ei:
        big = torch.zeros(1, 256, 101, 2, device="cuda")
        hot.forward(big, big, (3232, 64), (3232, 64))
    assert ei.value.code == cabi.OETR_E_SHAPE
    with pytest.raises(ValueError):
        hot.forward(torch.zeros(1, 128, 4, 4, device="cuda"), torch.zeros(1, 128, 4, 4, device="cuda"), (128, 128),
                    (128, 128))
    need = ctypes.c_size_t()
    lib = cabi.load_library()
    assert lib.oetr_workspace_bytes(hot._handle, 1, 4, 4, 4, 4, ctypes.byref(need)) == 0 and need.value > 0
    f = torch.zeros(1, 256, 4, 4, device="cuda")
    out = torch.zeros(1, 4, device="cuda")
    small = torch.empty(1024, dtype=torch.uint8, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.oetr_forward(hot._handle, p(f), p(f), 1, 4, 4, 4, 4, 128, 128, 128, 128, 1, p(out), p(out), None, None,
                          None, None, p(small), small.numel(), None)
    assert rc == cabi.OETR_E_NOMEM
    hot.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_full_size_properties(precision):
    """BASELINE config 2 (batch 32, 640x640 -> 20x20 maps): size-independent properties."""
    W = weights.synthetic_hot_path_weights(0)
    b = 32
    f1 = torch.from_numpy(weights.synthetic_features(b, 20, 20, seed=9, tag="p1")).cuda()
    f2 = torch.from_numpy(weights.synthetic_features(b, 20, 20, seed=9, tag="p2")).cuda()
    hot = oetr_b200.OverlapHotPath(W, precision=precision)
    a1, a2 = hot.forward(f1, f2, (640, 640), (640, 640), clamp=False)
    # determinism
    c1, c2 = hot.forward(f1, f2, (640, 640), (640, 640), clamp=False)
    assert torch.equal(a1, c1) and torch.equal(a2, c2)
    # pairs are independent: permuting the batch permutes the boxes, and a sub-batch reproduces its rows.  Not bit for
    # bit on the fp16 path: its flat encoder tiling cuts the concatenated images of a batch into 128-token tiles, so
    # the order in which the linear-attention partial sums of an image are added depends on its position in the batch
    # (fp32 rounding, ~1e-6 of the image side; tolerance 5e-3 px = 8e-6)
    perm = torch.randperm(b, generator=torch.Generator().manual_seed(0)).cuda()
    p1, p2 = hot.forward(f1[perm], f2[perm], (640, 640), (640, 640), clamp=False)
    assert torch.allclose(p1, a1[perm], rtol=0, atol=6e-2) and torch.allclose(p2, a2[perm], rtol=0, atol=6e-2)
    s1, s2 = hot.forward(f1[5:8], f2[5:8], (640, 640), (640, 640), clamp=False)
    assert torch.allclose(s1, a1[5:8], rtol=0, atol=6e-2) and torch.allclose(s2, a2[5:8], rtol=0, atol=6e-2)
    # image-size linearity of the box assembly: doubling the declared image size doubles stride and extents
    d1, _ = hot.forward(f1, f2, (1280, 1280), (1280, 1280), clamp=False)
    assert torch.allclose(d1, 2 * a1, rtol=1e-5, atol=1e-3)
    # oracle on a 4-pair sample of the same batch
    want = orc.hot_path(W, f1[:4].cpu().numpy(), f2[:4].cpu().numpy(), (640, 640), (640, 640))
    assert np.abs(a1[:4].cpu().numpy() - want["box1_raw"]).max() / 640 < TOL[precision]["box"]
    assert np.abs(a2[:4].cpu().numpy() - want["box2_raw"]).max() / 640 < TOL[precision]["box"]
    # host-buffer entry point gives the same boxes
    h1, h2 = hot.forward_host(f1.cpu().numpy(), f2.cpu().numpy(), (640, 640), (640, 640), clamp=False)
    assert np.array_equal(h1, a1.cpu().numpy()) and np.array_equal(h2, a2.cpu().numpy())
    hot.close()


def test_swap_symmetry():
    """With identical query embeddings the model is symmetric in its two inputs: swapping them swaps the boxes."""
    W = weights.synthetic_hot_path_weights(0)
    W["query_embed2.weight"] = W["query_embed1.weight"].copy()
    f1 = torch.from_numpy(weights.synthetic_features(2, 9, 12, seed=3, tag="s1")).cuda()
    f2 = torch.from_numpy(weights.synthetic_features(2, 10, 7, seed=3, tag="s2")).cuda()
    hot = oetr_b200.OverlapHotPath(W, precision="fp32")
    a1, a2 = hot.forward(f1, f2, (288, 384), (320, 224), clamp=False)
    b1, b2 = hot.forward(f2, f1, (320, 224), (288, 384), clamp=False)
    assert torch.allclose(a1, b2, rtol=0, atol=1e-3) and torch.allclose(a2, b1, rtol=0, atol=1e-3)
    hot.close()


def test_selftest_tcgen05_building_blocks():
    lib = cabi.load_library()
    errs = (ctypes.c_float * 16)()
    cabi.check(lib.oetr_selftest_tcgen05(errs, 16), lib)
    vals = list(errs)
    print("tcgen05 selftest errors:", vals)
    assert all(v == v and v < 2e-3 for v in vals[:8]), vals


def test_geometry_changes_between_calls_and_host_entry():
    """One handle, changing feature-map geometries call after call (the handle caches the tile-blocked position
    rows per geometry), device and host-buffer entry points, batch 1 and an odd batch."""
    W = weights.synthetic_hot_path_weights(0)
    hot = oetr_b200.OverlapHotPath(W, precision="fp16")
    for b, fm1, fm2 in ((1, (20, 20), (20, 20)), (3, (15, 20), (20, 15)), (1, (20, 20), (20, 20)), (5, (26, 26), (13, 9))):
        hw1, hw2 = (fm1[0] * 32, fm1[1] * 32), (fm2[0] * 32, fm2[1] * 32)
        f1 = weights.synthetic_features(b, *fm1, seed=17, tag="g1")
        f2 = weights.synthetic_features(b, *fm2, seed=17, tag="g2")
        want = orc.hot_path(W, f1, f2, hw1, hw2)
        a1, a2 = hot.forward(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), hw1, hw2, clamp=False)
        h1, h2 = hot.forward_host(f1, f2, hw1, hw2, clamp=False)
        hot.poll_error()
        for got, ref, hw in ((a1.cpu().numpy(), want["box1_raw"], hw1), (a2.cpu().numpy(), want["box2_raw"], hw2),
                             (h1, want["box1_raw"], hw1), (h2, want["box2_raw"], hw2)):
            assert np.abs(got - ref).max() / max(hw) < TOL["fp16"]["box"], (fm1, fm2)
    hot.close()


def test_large_token_regime_840():
    """BASELINE config 4: 840x840 pairs, batch 16 (26x26 maps, 676 tokens = 6 tiles per image)."""
    W = weights.synthetic_hot_path_weights(0)
    b = 16
    f1 = torch.from_numpy(weights.synthetic_features(b, 26, 26, seed=4, tag="l1")).cuda()
    f2 = torch.from_numpy(weights.synthetic_features(b, 26, 26, seed=4, tag="l2")).cuda()
    hot = oetr_b200.OverlapHotPath(W, precision="fp16")
    a1, a2 = hot.forward(f1, f2, (840, 840), (840, 840), clamp=False)
    c1, c2 = hot.forward(f1, f2, (840, 840), (840, 840), clamp=False)
    assert torch.equal(a1, c1) and torch.equal(a2, c2)
    want = orc.hot_path(W, f1[:2].cpu().numpy(), f2[:2].cpu().numpy(), (840, 840), (840, 840))
    assert np.abs(a1[:2].cpu().numpy() - want["box1_raw"]).max() / 840 < TOL["fp16"]["box"]
    assert np.abs(a2[:2].cpu().numpy() - want["box2_raw"]).max() / 840 < TOL["fp16"]["box"]
    hot.poll_error()
    hot.close()


def test_sub_batch_scheduling_is_invisible():
    """The fp16 path cuts a batch into sub-batches on handle-owned streams (oetr_set_chunk_pairs).  Pairs are
    independent and every kernel is deterministic, so any split gives the same boxes -- up to the fp32 summation order
    of the flat encoder tiling, which depends on the composition of a (sub-)batch: a last-bit change of a summary can
    flip the fp16 rounding of a single-term operand (q / k projections, precision map of DESIGN.md section 3), so the
    bound is the map's noise level, 1e-4 of the image side (6e-2 px), not fp32 round-off; the same split is bit-identical on the device entry, the host-buffer entry and any stream (uneven splits,
    more sub-batches than the cap)."""
    W = weights.synthetic_hot_path_weights(0)
    b = 21
    n1 = weights.synthetic_features(b, 20, 20, seed=31, tag="c1")
    n2 = weights.synthetic_features(b, 14, 17, seed=31, tag="c2")
    f1, f2 = torch.from_numpy(n1).cuda(), torch.from_numpy(n2).cuda()
    hw1, hw2 = (640, 640), (448, 544)
    hot = oetr_b200.OverlapHotPath(W, precision="fp16")
    hot.set_chunk_pairs(0)
    hot.forward(f1, f2, hw1, hw2, clamp=False)             # first call of a geometry also builds its position rows
    r1, r2 = hot.forward(f1, f2, hw1, hw2, clamp=False)
    per_forward = hot.last_launch_count
    assert per_forward in (25, 26)                         # 26: flat encoder tiling adds the re-tiling kernel
    want = orc.hot_path(W, n1[:2], n2[:2], hw1, hw2)
    assert np.abs(r1[:2].cpu().numpy() - want["box1_raw"]).max() / 640 < TOL["fp16"]["box"]
    side = torch.cuda.Stream()
    for pairs, chunks in ((8, 3), (5, 5), (2, 8), (1, 8), (20, 2), (21, 1), (64, 1)):
        hot.set_chunk_pairs(pairs)
        a1, a2 = hot.forward(f1, f2, hw1, hw2, clamp=False)
        assert hot.last_launch_count == per_forward * chunks, (pairs, hot.last_launch_count)
        assert torch.allclose(a1, r1, rtol=0, atol=6e-2) and torch.allclose(a2, r2, rtol=0, atol=6e-2), pairs
        h1, h2 = hot.forward_host(n1, n2, hw1, hw2, clamp=False)
        assert np.array_equal(h1, a1.cpu().numpy()) and np.array_equal(h2, a2.cpu().numpy()), pairs
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            s1, s2 = hot.forward(f1, f2, hw1, hw2, clamp=False)
        side.synchronize()
        assert torch.equal(s1, a1) and torch.equal(s2, a2), pairs
    hot.poll_error()
    hot.close()


def test_host_requests_in_flight():
    """oetr_forward_host_submit / _wait: four requests of different geometry in flight, results equal to the
    stream-ordered entry; a fifth submit is refused until a ticket has been waited for; a request repeated on the
    same slot is captured into a CUDA graph and replayed with identical results."""
    W = weights.synthetic_hot_path_weights(0)
    hot = oetr_b200.OverlapHotPath(W, precision="fp16")
    reqs = []
    for b, fm1, fm2 in ((12, (20, 20), (20, 20)), (3, (9, 11), (26, 26)), (1, (20, 20), (20, 20)), (12, (20, 20), (20, 20)),
                        (20, (20, 20), (15, 20))):
        hw1, hw2 = (fm1[0] * 32, fm1[1] * 32), (fm2[0] * 32, fm2[1] * 32)
        f1 = weights.synthetic_features(b, *fm1, seed=23, tag="q1")
        f2 = weights.synthetic_features(b, *fm2, seed=23, tag="q2")
        r1, r2 = hot.forward(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), hw1, hw2, clamp=False)
        reqs.append((f1, f2, hw1, hw2, r1.cpu().numpy(), r2.cpu().numpy()))
    t = [hot.submit_host(*reqs[i][:4], clamp=False) for i in range(4)]
    with pytest.raises(cabi.OetrError):
        hot.submit_host(*reqs[4][:4], clamp=False)
    a = hot.wait_host(t[0])
    t4 = hot.submit_host(*reqs[4][:4], clamp=False)
    got = [a] + [hot.wait_host(x) for x in t[1:]] + [hot.wait_host(t4)]
    for g_, req in zip(got, reqs):
        assert np.array_equal(g_[0], req[4]) and np.array_equal(g_[1], req[5])      # same split, same arithmetic
    with pytest.raises(cabi.OetrError):
        cabi.check(hot._lib.oetr_forward_host_wait(hot._handle, 12345, None, None), hot._lib)
    # the same request over and over: every slot sees it three times (eager, capture + replay, replay)
    req = reqs[4]
    tickets = []
    for i in range(14):
        tickets.append(hot.submit_host(*req[:4], clamp=False))
        if len(tickets) == 3:
            g_ = hot.wait_host(tickets.pop(0))
            assert np.array_equal(g_[0], req[4]) and np.array_equal(g_[1], req[5]), i
    while tickets:
        g_ = hot.wait_host(tickets.pop(0))
        assert np.array_equal(g_[0], req[4]) and np.array_equal(g_[1], req[5])
    hot.poll_error()
    hot.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name", sorted(MASK_CASES))
def test_masked_parity_with_oracle_and_golden(name, precision):
    """Float padding masks on both images (oetr_forward_masked): stage parity against the oracle and the committed
    outputs of the reference run with the same masks; sub-batches slice the masks like the features."""
    W, f1, f2, (b, fm1, fm2, hw1, hw2, attention, _, _), g = load_case(name)
    m1, m2 = load_masks(name)
    want = orc.hot_path(W, f1, f2, hw1, hw2, mask1=m1, mask2=m2)
    tol = TOL[precision]
    hot = oetr_b200.OverlapHotPath(W, precision=precision)
    t = lambda a: torch.from_numpy(a).cuda()
    for pairs in (0, 1):
        hot.set_chunk_pairs(pairs)
        b1, b2, dbg = hot.forward(t(f1), t(f2), hw1, hw2, clamp=False, debug=(pairs == 0), mask1=t(m1), mask2=t(m2)) \
            if pairs == 0 else hot.forward(t(f1), t(f2), hw1, hw2, clamp=False, mask1=t(m1), mask2=t(m2)) + (None,)
        for i, (got, hw) in enumerate(((b1, hw1), (b2, hw2)), 1):
            assert np.abs(got.cpu().numpy() - want["box%d_raw" % i]).max() / max(hw) < tol["box"], (pairs, i)
            assert np.abs(got.cpu().numpy() - g["box%d_raw" % i]).max() / max(hw) < tol["box"] + 2e-5, (pairs, i)
        if dbg is not None:
            for k in ("memory1", "memory2", "hs1", "hs2", "tlbr1", "tlbr2"):
                assert rel_err(dbg[k].cpu().numpy(), want[k]) < tol["mid"], k
    hot.poll_error()
    with pytest.raises(ValueError):
        hot.forward(t(f1), t(f2), hw1, hw2, mask1=t(m1))
    hot.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name", sorted(STRESS_CASES))
def test_stress_scales_against_reference_outputs(name, precision):
    """Trained-like scales and the full batch of BASELINE config 2, against the committed outputs of the real reference
    (tests/golden/make_golden.py --stress-only): features x3 / x0.1, LayerNorm gains up to 3, PyTorch-default init of
    tlbr_reg / heatmap_conv, and all 32 pairs of a 640x640 batch.  The precision map of the fp16 path
    was chosen on these scales; the bar stays 1e-3 of the image side."""
    W, f1, f2, c, g = load_stress_case(name)
    got = _run(W, f1, f2, c["hw1"], c["hw2"], "linear", precision, clamp=False)
    tol = TOL[precision]
    for i, hw in ((1, c["hw1"]), (2, c["hw2"])):
        side = max(hw)
        assert np.abs(got["box%d" % i] - g["box%d_raw_f64" % i]).max() / side < tol["box"], (i, np.abs(got["box%d" % i] - g["box%d_raw_f64" % i]).max() / side)
        assert np.abs(got["cxy%d" % i] - g["cxy%d_f64" % i]).max() / side < tol["box"]
        assert rel_err(got["tlbr%d" % i], g["tlbr%d_f64" % i]) < tol["mid"]
        assert rel_err(got["hs%d" % i], g["hs%d_f64" % i]) < tol["mid"]
    clamped = _run(W, f1, f2, c["hw1"], c["hw2"], "linear", precision, clamp=True)
    for i, hw in ((1, c["hw1"]), (2, c["hw2"])):
        assert np.abs(clamped["box%d" % i] - g["box%d" % i]).max() / max(hw) < tol["box"] + 2e-5


def test_full_attention_tensor_core_path_properties():
    """attention_mode='full' on the tcgen05 path (k_proj_mlp + k_attn): agreement with the fp32 CUDA-core path and the
    oracle on shapes the golden files do not cover (key counts that are not a multiple of the 64-key chunk, more than
    one key tile, ragged pairs, batch > 1), and determinism."""
    W = weights.synthetic_hot_path_weights(0)
    for b, fm1, fm2 in ((2, (9, 7), (5, 13)), (3, (13, 11), (20, 20)), (1, (26, 26), (4, 4))):
        hw1, hw2 = (fm1[0] * 32, fm1[1] * 32), (fm2[0] * 32, fm2[1] * 32)
        f1 = weights.synthetic_features(b, *fm1, seed=6, tag="fa1")
        f2 = weights.synthetic_features(b, *fm2, seed=6, tag="fa2")
        want = orc.hot_path(W, f1, f2, hw1, hw2, attention="full")
        got = _run(W, f1, f2, hw1, hw2, "full", "fp16", clamp=False)
        again = _run(W, f1, f2, hw1, hw2, "full", "fp16", clamp=False)
        for i, hw in ((1, hw1), (2, hw2)):
            err = np.abs(got["box%d" % i] - want["box%d_raw" % i]).max() / max(hw)
            assert err < TOL["fp16"]["box"], (fm1, fm2, i, err)
            assert rel_err(got["memory%d" % i], want["memory%d" % i]) < TOL["fp16"]["mid"]
            assert np.array_equal(got["box%d" % i], again["box%d" % i])
