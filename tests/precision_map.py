"""Per-contraction operand-precision experiment for the tcgen05 path -- TEST INFRASTRUCTURE (CPU only, numpy).

Models the DATAFLOW OF THE KERNELS (folded merge weights M_img = blockdiag(KV/S) Wm^T, phi(q)/Z as the A operand,
K^T V with tokens as the contraction dimension, decoder K/V summaries, 9-tap heat-map convolution) in fp64 with a
selectable rounding of every GEMM operand:
    "x"  exact (stands for the hi+lo split: 22 bits, indistinguishable from fp32 here)
    "h"  one fp16 value (the lo part is dropped -> one MMA term less)
A GEMM with operands (a, w) costs 1 + [a split] + [w split] MMA terms.  The script measures the box error
(max |box - box_exact| / image side, unclamped boxes) each single "h" choice causes on a set of cases (golden
geometries + stress scales) and evaluates candidate maps.  Run:   python tests/precision_map.py [--maps]
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oetr_oracle as orc  # noqa: E402
from oetr_b200 import weights  # noqa: E402

SITES = ["a_q", "w_q", "a_kv", "w_k", "w_v", "a_kf", "a_v", "a_m", "w_m", "a_1", "w_1", "a_2", "w_2",
         "a_dk", "a_dv", "w_dk", "w_dv", "a_dkf", "a_dvv", "a_c", "w_c"]


def f16(a):
    return np.asarray(a).astype(np.float16).astype(np.float64)


class Map(dict):
    def r(self, site, a):
        return f16(a) if self.get(site, "x") == "h" else a


def enc_layer(W, pre, x, s, xp, sp, m):
    g = lambda n: W[pre + n]
    S = s.shape[1]
    qi = orc.layer_norm(x, g("pre_norm_q.weight"), g("pre_norm_q.bias")) + xp
    kvi = orc.layer_norm(s, g("pre_norm_kv.weight"), g("pre_norm_kv.bias")) + sp
    q = m.r("a_q", qi) @ m.r("w_q", g("q_proj.weight")).T
    k = m.r("a_kv", kvi) @ m.r("w_k", g("k_proj.weight")).T
    v = m.r("a_kv", kvi) @ m.r("w_v", g("v_proj.weight")).T
    n = x.shape[0]
    Kf = orc.elu_feature_map(k).reshape(n, S, 8, 32)
    Q = orc.elu_feature_map(q).reshape(n, -1, 8, 32)
    KV = np.einsum("nshd,nshv->nhdv", m.r("a_kf", Kf), m.r("a_v", v.reshape(n, S, 8, 32))) / S
    Ks = Kf.sum(axis=1) / S
    Z = 1.0 / (np.einsum("nlhd,nhd->nlh", Q, Ks) + orc.ATTN_EPS / S)
    A = (Q * Z[..., None]).reshape(n, -1, 256)
    Wm = g("merge.weight")                                   # [n_out, h*32+e]
    # M_img[n_out, h*32+d] = sum_e Wm[n_out, h*32+e] KV[h][d][e]
    M = np.einsum("ohe,nhde->nohd", Wm.reshape(256, 8, 32), KV).reshape(n, 256, 256)
    msg = np.einsum("nlk,nok->nlo", m.r("a_m", A), m.r("w_m", M))
    x = x + msg
    h = orc.gelu_erf(m.r("a_1", orc.layer_norm(x, g("norm2.weight"), g("norm2.bias"))) @ m.r("w_1", g("mlp.0.weight")).T)
    return x + m.r("a_2", h) @ m.r("w_2", g("mlp.2.weight")).T


def dec_layer(W, pre, tgt, mem, qe, mp, m):
    g = lambda n: W[pre + n]
    t2 = orc.layer_norm(tgt, g("norm1.weight"), g("norm1.bias"))
    qk = t2 + qe
    tgt = tgt + orc._mha(W, pre + "self_attn.", qk, qk, t2)
    t2 = orc.layer_norm(tgt, g("norm2.weight"), g("norm2.bias"))
    p = pre + "multihead_attn."
    q = (t2 + qe) @ W[p + "q_proj.weight"].T + W[p + "q_proj.bias"]
    k = m.r("a_dk", mem + mp) @ m.r("w_dk", W[p + "k_proj.weight"]).T + W[p + "k_proj.bias"]
    v = m.r("a_dv", mem) @ m.r("w_dv", W[p + "v_proj.weight"]).T + W[p + "v_proj.bias"]
    n, S = mem.shape[:2]
    Kf = orc.elu_feature_map(k).reshape(n, S, 8, 32)
    Q = orc.elu_feature_map(q).reshape(n, 1, 8, 32)
    KV = np.einsum("nshd,nshv->nhdv", m.r("a_dkf", Kf), m.r("a_dvv", v.reshape(n, S, 8, 32)))
    Z = 1.0 / (np.einsum("nlhd,nhd->nlh", Q, Kf.sum(axis=1)) + orc.ATTN_EPS)
    o = np.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z).reshape(n, 1, 256)
    tgt = tgt + o @ W[p + "merge.weight"].T
    t2 = orc.layer_norm(tgt, g("norm3.weight"), g("norm3.bias"))
    return tgt + np.maximum(t2 @ g("mlp.0.weight").T, 0.0) @ g("mlp.2.weight").T


def center(W, hs, mem, hf, wf, img_h, m):
    n = mem.shape[0]
    att = np.einsum("blc,bnc->bln", mem, hs)
    heat = m.r("a_c", mem * att).transpose(0, 2, 1).reshape(n, 256, hf, wf)
    y = orc.conv3x3_same(heat, m.r("w_c", W["heatmap_conv.0.weight"]), W["heatmap_conv.0.bias"])
    y = np.maximum(orc.group_norm(y, W["heatmap_conv.1.weight"], W["heatmap_conv.1.bias"]), 0.0)
    z = np.einsum("nchw,c->nhw", y, W["heatmap_conv.3.weight"].reshape(-1)) + W["heatmap_conv.3.bias"][0]
    z = z.reshape(n, hf * wf)
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p = p / p.sum(axis=1, keepdims=True)
    stride = img_h // hf
    ys, xs = np.meshgrid(np.arange(hf), np.arange(wf), indexing="ij")
    return np.stack([(p * (xs.reshape(-1) + 0.5) * stride).sum(axis=1), (p * (ys.reshape(-1) + 0.5) * stride).sum(axis=1)], axis=1)


def run(W, f1, f2, hw1, hw2, m):
    n = f1.shape[0]
    hf1, wf1 = f1.shape[2:]
    hf2, wf2 = f2.shape[2:]
    pe = orc.pe_table()
    p0 = pe[:, :hf1, :wf1].reshape(256, -1).T
    p1 = pe[:, :hf2, :wf2].reshape(256, -1).T
    x0 = f1.reshape(n, 256, -1).transpose(0, 2, 1)
    x1 = f2.reshape(n, 256, -1).transpose(0, 2, 1)
    for i in range(8):
        pre = "transformer.encoder.%d." % i
        if i % 2 == 0:
            x0, x1 = enc_layer(W, pre, x0, x0, p0, p0, m), enc_layer(W, pre, x1, x1, p1, p1, m)
        else:
            x0, x1 = enc_layer(W, pre, x0, x1, p0, p1, m), enc_layer(W, pre, x1, x0, p1, p0, m)
    boxes = []
    for mem, qe, pos, hw, (hf, wf) in ((x0, W["query_embed1.weight"], p0, hw1, (hf1, wf1)),
                                       (x1, W["query_embed2.weight"], p1, hw2, (hf2, wf2))):
        t = np.zeros((n, 1, 256))
        qe = np.broadcast_to(qe[None], (n, 1, 256))
        for j in range(2):
            t = dec_layer(W, "transformer.decoder.layers.%d." % j, t, mem, qe, pos, m)
        cxy = center(W, t, mem, hf, wf, hw[0], m)
        tlbr = orc.size_regression(W, t)
        boxes.append(orc.box_tlbr_to_xyxy(cxy, tlbr, hw[0], hw[1], False) / max(hw))
    return boxes


def make_cases():
    """(name, weights, f1, f2, hw1, hw2): golden geometries + stress scales (features x3 / x0.1, LN gains up to 3,
    PyTorch-default init of tlbr_reg / heatmap_conv)."""
    out = []
    W0 = {k: v.astype(np.float64) for k, v in weights.synthetic_hot_path_weights(0).items()}
    W3 = {k: v.astype(np.float64) for k, v in weights.synthetic_hot_path_weights(3).items()}
    def feats(b, fm1, fm2, seed, scale=1.0):
        return (weights.synthetic_features(b, *fm1, seed=seed, tag="feat1").astype(np.float64) * scale,
                weights.synthetic_features(b, *fm2, seed=seed, tag="feat2").astype(np.float64) * scale)
    out.append(("b2_640", W0, *feats(2, (20, 20), (20, 20), 1), (640, 640), (640, 640)))
    out.append(("ragged", W0, *feats(1, (20, 20), (15, 20), 2), (640, 640), (480, 640)))
    out.append(("b1_840", W0, *feats(1, (26, 26), (26, 26), 3), (840, 840), (840, 840)))
    out.append(("tiny_b3", W3, *feats(3, (5, 7), (4, 6), 4), (160, 224), (128, 192)))
    out.append(("feat_x3", W0, *feats(2, (20, 20), (20, 20), 11, 3.0), (640, 640), (640, 640)))
    out.append(("feat_x0.1", W0, *feats(2, (20, 20), (20, 20), 12, 0.1), (640, 640), (640, 640)))
    Wg = dict(W0)
    rng = np.random.default_rng(5)
    for k in Wg:
        if "norm" in k and k.endswith("weight") and "heatmap" not in k:
            Wg[k] = W0[k] * rng.uniform(0.5, 3.0, size=W0[k].shape)
    out.append(("ln_gain3", Wg, *feats(2, (20, 20), (20, 20), 13), (640, 640), (640, 640)))
    return out


def err(boxes, ref):
    return max(np.abs(b - r).max() for b, r in zip(boxes, ref))


def main():
    cases = make_cases()
    refs = [run(W, f1, f2, hw1, hw2, Map()) for (_, W, f1, f2, hw1, hw2) in cases]
    names = [c[0] for c in cases]
    print("%-8s" % "site", " ".join("%9s" % n for n in names), "      max")
    single = {}
    for s in ([] if "--maps-only" in sys.argv else SITES):
        es = [err(run(W, f1, f2, hw1, hw2, Map({s: "h"})), r) for (_, W, f1, f2, hw1, hw2), r in zip(cases, refs)]
        single[s] = es
        print("%-8s" % s, " ".join("%9.2e" % e for e in es), "%9.2e" % max(es))
    if "--maps-only" not in sys.argv:
        allh = Map({s: "h" for s in SITES})
        es = [err(run(W, f1, f2, hw1, hw2, allh), r) for (_, W, f1, f2, hw1, hw2), r in zip(cases, refs)]
        print("%-8s" % "ALL h", " ".join("%9.2e" % e for e in es), "%9.2e" % max(es))
    if "--maps" in sys.argv or "--maps-only" in sys.argv:
        for label, hs in CANDIDATES.items():
            mp = Map({s: "h" for s in hs})
            es = [err(run(W, f1, f2, hw1, hw2, mp), r) for (_, W, f1, f2, hw1, hw2), r in zip(cases, refs)]
            print("%-28s" % label, " ".join("%9.2e" % e for e in es), "%9.2e" % max(es))


CANDIDATES = {
    # the map adopted in round 2 (k_enc / decoder K,V launches): q GEMM 1 term, k GEMM 2 terms (activation split),
    # decoder k 1 term, decoder v 2 terms (weight split); everything else stays a 3-term product
    "round-2 map": ["a_q", "w_q", "w_k", "a_dk", "w_dk", "a_dv"],
    "round-2 map + a_kv": ["a_q", "w_q", "w_k", "a_dk", "w_dk", "a_dv", "a_kv"],
    "round-2 map + dec KV h": ["a_q", "w_q", "w_k", "a_dk", "w_dk", "a_dv", "a_dkf", "a_dvv"],
    "weights h, acts x": [s for s in SITES if s.startswith("w_")],
    "acts h, weights x": [s for s in SITES if s.startswith("a_")],
}

if __name__ == "__main__":
    main()
