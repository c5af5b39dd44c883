"""CPU oracle for the OETR hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement (fp64 by default) of the reference's pair-wise feature-correlation transformer and
overlap-box head.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this file; the product path (imagematching-oetr_b200/) never does and fails loudly when its CUDA
library is missing.

Parity status: the reference ships no golden vectors or tests for this path, so the
oracle is pinned against outputs of the UNMODIFIED reference modules run in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py checks every stage).

Each function cites the reference lines it follows (paths relative to the reference root).
Weights are a dict {state_dict key -> ndarray}, the hot-path subset listed in SURVEY.md section 8(a)-W.
"""
import math

import numpy as np

try:  # scipy is present in the image; keep a slow fallback so the oracle never silently changes maths
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

D_MODEL = 256
NHEAD = 8
HEAD_DIM = 32
N_ENCODER = 8
N_DECODER = 2
LN_EPS = 1e-5
ATTN_EPS = 1e-6


# --------------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------------
def pe_table(max_shape=(100, 100), d_model=D_MODEL, dtype=np.float64):
    """PositionEncodingSine buffer, src/models/utils.py:185-198, including its operator-precedence quirk:
    `(-math.log(10000.0) / d_model // 2)` is floor((-ln 1e4 / 256) / 2) = -1.0, so div_term = exp(-[0,2,..,126])."""
    h, w = max_shape
    y_pos = np.cumsum(np.ones((h, w), dtype=np.float32), axis=0)[None]      # starts at 1
    x_pos = np.cumsum(np.ones((h, w), dtype=np.float32), axis=1)[None]
    factor = (-math.log(10000.0) / d_model // 2)                             # == -1.0
    div = np.exp(np.arange(0, d_model // 2, 2, dtype=np.float32) * np.float32(factor)).astype(np.float32)
    div = div[:, None, None]
    pe = np.zeros((d_model, h, w), dtype=np.float32)
    # the reference evaluates sin/cos in fp32 on fp32 products
    pe[0::4] = np.sin((x_pos * div).astype(np.float32))
    pe[1::4] = np.cos((x_pos * div).astype(np.float32))
    pe[2::4] = np.sin((y_pos * div).astype(np.float32))
    pe[3::4] = np.cos((y_pos * div).astype(np.float32))
    return pe.astype(dtype)


def layer_norm(x, gamma, beta, eps=LN_EPS):
    """nn.LayerNorm over the last dim (biased variance)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma + beta


def elu_feature_map(x):
    """src/models/linear_attention.py:12-13: elu(x)+1."""
    return np.where(x > 0, x + 1.0, np.exp(np.minimum(x, 0.0)))


def gelu_erf(x):
    """nn.GELU() default (exact erf form), src/models/transformer.py:93."""
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def linear_attention(q, k, v, eps=ATTN_EPS, q_mask=None, kv_mask=None):
    """LinearAttention.forward, src/models/linear_attention.py:22-50.
    q [N,L,H,D]; k,v [N,S,H,D] -> [N,L,H,D]; q_mask [N,L], kv_mask [N,S] (float, optional: :36-41)."""
    Q = elu_feature_map(q)
    K = elu_feature_map(k)
    if q_mask is not None:
        Q = Q * q_mask[:, :, None, None]
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None]
        v = v * kv_mask[:, :, None, None]
    s_len = v.shape[1]
    v = v / s_len                                                        # :43-44
    KV = np.einsum("nshd,nshv->nhdv", K, v)                              # :45
    Z = 1.0 / (np.einsum("nlhd,nhd->nlh", Q, K.sum(axis=1)) + eps)       # :46
    return np.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * s_len            # :47-48


def full_attention(q, k, v):
    """FullAttention.forward, src/models/linear_attention.py:59-87 (masks None, dropout off)."""
    QK = np.einsum("nlhd,nshd->nlsh", q, k)
    temp = 1.0 / q.shape[3] ** 0.5
    A = temp * QK
    A = A - A.max(axis=2, keepdims=True)
    A = np.exp(A)
    A = A / A.sum(axis=2, keepdims=True)
    return np.einsum("nlsh,nshd->nlhd", A, v)


def _heads(x):
    n, l, _ = x.shape
    return x.reshape(n, l, NHEAD, HEAD_DIM)


def _identity(a):
    return a


def round_fp16(a):
    """Operand-rounding model of the tcgen05 path (fp16 operands, wide accumulation); tests use it to bound the
    error the FP16 path may show.  Not part of the reference."""
    return np.asarray(a).astype(np.float16).astype(np.asarray(a).dtype)


def linear_attention_quant(q, k, v, rnd, eps=ATTN_EPS, q_mask=None, kv_mask=None):
    """linear_attention with the operand rounding points of the tensor-core path: elu(k)+1, v and elu(q)+1 are
    rounded before the K^T V and Q.KV contractions; KV is rounded as the B operand; Ksum / Z stay wide."""
    Q = elu_feature_map(q)
    K = elu_feature_map(k)
    if q_mask is not None:
        Q = Q * q_mask[:, :, None, None]
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None]
        v = v * kv_mask[:, :, None, None]
    K = rnd(K)
    KV = rnd(np.einsum("nshd,nshv->nhdv", K, rnd(v)))
    Z = 1.0 / (np.einsum("nlhd,nhd->nlh", Q, K.sum(axis=1)) + eps)
    return np.einsum("nlhd,nhdv,nlh->nlhv", rnd(Q), KV, Z)


def encoder_layer(W, prefix, x, source, x_pos, s_pos, attention="linear", rnd=_identity, x_mask=None, s_mask=None):
    """EncoderLayer.forward, src/models/transformer.py:104-142.  x [N,L,C], source [N,S,C], pos [L,C]/[S,C].
    rnd: operand rounding model (identity = the reference's arithmetic)."""
    g = lambda name: W[prefix + name]
    query = layer_norm(x, g("pre_norm_q.weight"), g("pre_norm_q.bias")) + x_pos           # :118,:121-124
    kv = layer_norm(source, g("pre_norm_kv.weight"), g("pre_norm_kv.bias")) + s_pos        # key == value input
    q = _heads(rnd(query) @ rnd(g("q_proj.weight")).T)
    k = _heads(rnd(kv) @ rnd(g("k_proj.weight")).T)
    v = _heads(rnd(kv) @ rnd(g("v_proj.weight")).T)
    if attention != "linear":
        if x_mask is not None or s_mask is not None:
            raise NotImplementedError("masks are restated for linear attention only")
        att = full_attention(q, k, v)
    elif rnd is _identity:
        att = linear_attention(q, k, v, q_mask=x_mask, kv_mask=s_mask)
    else:
        att = linear_attention_quant(q, k, v, rnd, q_mask=x_mask, kv_mask=s_mask)
    msg = rnd(att.reshape(x.shape)) @ rnd(g("merge.weight")).T                             # :137
    x = x + msg                                                                            # :140
    h = gelu_erf(rnd(layer_norm(x, g("norm2.weight"), g("norm2.bias"))) @ rnd(g("mlp.0.weight")).T)
    return x + rnd(h) @ rnd(g("mlp.2.weight")).T                                           # :141-142


def _mha(W, prefix, q_in, k_in, v_in, kv_mask=None):
    """MultiHeadAttention.forward (always linear attention, biased projections), transformer.py:55-72."""
    q = _heads(q_in @ W[prefix + "q_proj.weight"].T + W[prefix + "q_proj.bias"])
    k = _heads(k_in @ W[prefix + "k_proj.weight"].T + W[prefix + "k_proj.bias"])
    v = _heads(v_in @ W[prefix + "v_proj.weight"].T + W[prefix + "v_proj.bias"])
    out = linear_attention(q, k, v, kv_mask=kv_mask)
    return out.reshape(q_in.shape[0], q_in.shape[1], D_MODEL) @ W[prefix + "merge.weight"].T


def decoder_layer(W, prefix, tgt, memory, tgt_pos, m_pos, memory_mask=None):
    """DecoderLayer.forward, src/models/transformer.py:224-255 (dropout = identity in eval)."""
    g = lambda name: W[prefix + name]
    t2 = layer_norm(tgt, g("norm1.weight"), g("norm1.bias"))
    qk = t2 + tgt_pos
    tgt = tgt + _mha(W, prefix + "self_attn.", qk, qk, t2)
    t2 = layer_norm(tgt, g("norm2.weight"), g("norm2.bias"))
    tgt = tgt + _mha(W, prefix + "multihead_attn.", t2 + tgt_pos, memory + m_pos, memory, memory_mask)  # no LN, no pos on v
    t2 = layer_norm(tgt, g("norm3.weight"), g("norm3.bias"))
    t2 = np.maximum(t2 @ g("mlp.0.weight").T, 0.0) @ g("mlp.2.weight").T
    return tgt + t2


def query_transformer(W, feat0, feat1, pos0, pos1, attention="linear", prefix="transformer.",
                      return_layers=False, rnd=_identity, mask0=None, mask1=None):
    """QueryTransformer.forward, src/models/transformer.py:313-383.
    feat* [N,C,h,w] NCHW; pos* [C,h,w].  Returns hs0,hs1 [N,1,C], memory0,memory1 [N,L,C]."""
    n = feat0.shape[0]
    x0 = feat0.reshape(n, D_MODEL, -1).transpose(0, 2, 1)
    x1 = feat1.reshape(n, D_MODEL, -1).transpose(0, 2, 1)
    p0 = pos0.reshape(D_MODEL, -1).T
    p1 = pos1.reshape(D_MODEL, -1).T
    m0 = None if mask0 is None else np.asarray(mask0, dtype=x0.dtype).reshape(n, -1)          # transformer.py:341-344
    m1 = None if mask1 is None else np.asarray(mask1, dtype=x1.dtype).reshape(n, -1)
    layers = []
    for i in range(N_ENCODER):
        pre = "%sencoder.%d." % (prefix, i)
        if i % 2 == 0:                                                     # 'self'
            x0 = encoder_layer(W, pre, x0, x0, p0, p0, attention, rnd, m0, m0)
            x1 = encoder_layer(W, pre, x1, x1, p1, p1, attention, rnd, m1, m1)
        else:                                                              # 'cross': both read OLD partner
            s0, s1 = x1, x0
            x0 = encoder_layer(W, pre, x0, s0, p0, p1, attention, rnd, m0, m1)
            x1 = encoder_layer(W, pre, x1, s1, p1, p0, attention, rnd, m1, m0)
        if return_layers:
            layers.append((x0.copy(), x1.copy()))
    qe0 = np.broadcast_to(W["query_embed1.weight"][None], (n, 1, D_MODEL))
    qe1 = np.broadcast_to(W["query_embed2.weight"][None], (n, 1, D_MODEL))
    hs = []
    for mem, qe, pos, mm in ((x0, qe0, p0, m0), (x1, qe1, p1, m1)):
        t = np.zeros((n, 1, D_MODEL), dtype=mem.dtype)
        for j in range(N_DECODER):
            t = decoder_layer(W, "%sdecoder.layers.%d." % (prefix, j), t, mem, qe, pos, mm)
        hs.append(t)
    if return_layers:
        return hs[0], hs[1], x0, x1, layers
    return hs[0], hs[1], x0, x1


def group_norm(x, gamma, beta, groups=32, eps=1e-5):
    """nn.GroupNorm over [N,C,H,W]."""
    n, c, h, w = x.shape
    xg = x.reshape(n, groups, -1)
    mu = xg.mean(axis=2, keepdims=True)
    var = ((xg - mu) ** 2).mean(axis=2, keepdims=True)
    xg = (xg - mu) / np.sqrt(var + eps)
    return xg.reshape(n, c, h, w) * gamma[None, :, None, None] + beta[None, :, None, None]


def conv3x3_same(x, weight, bias):
    """nn.Conv2d(C,C,3,padding=1) on [N,C,H,W] via 9 shifted matmuls."""
    n, c, h, w = x.shape
    xp = np.zeros((n, c, h + 2, w + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, weight.shape[0], h, w), dtype=x.dtype)
    for dy in range(3):
        for dx in range(3):
            out += np.einsum("nchw,oc->nohw", xp[:, :, dy:dy + h, dx:dx + w], weight[:, :, dy, dx])
    return out + bias[None, :, None, None]


def center_estimation(W, hs, memory, hf, wf, img_h, mask=None):
    """OETR.center_estimation for one image, src/model.py:145-186 (softmax_temperature 1; mask [N,hf,wf]: logits of
    positions where mask.bool() is False are filled with -1e9, :167-171)."""
    n = memory.shape[0]
    att = np.einsum("blc,bnc->bln", memory, hs)                                             # :147
    heat = (memory * att).transpose(0, 2, 1).reshape(n, D_MODEL, hf, wf)                    # :152-155
    y = conv3x3_same(heat, W["heatmap_conv.0.weight"], W["heatmap_conv.0.bias"])
    y = np.maximum(group_norm(y, W["heatmap_conv.1.weight"], W["heatmap_conv.1.bias"]), 0.0)
    z = np.einsum("nchw,c->nhw", y, W["heatmap_conv.3.weight"].reshape(-1)) + W["heatmap_conv.3.bias"][0]
    z = z.reshape(n, hf * wf)
    if mask is not None:
        z = np.where(np.asarray(mask).reshape(n, hf * wf) != 0, z, -1e9)
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p = p / p.sum(axis=1, keepdims=True)                                                    # :173
    stride = img_h // hf                                                                    # :177 (h for both axes)
    ys, xs = np.meshgrid(np.arange(hf), np.arange(wf), indexing="ij")
    gx = (xs.reshape(-1) + 0.5) * stride                                                    # :103-107
    gy = (ys.reshape(-1) + 0.5) * stride
    return np.stack([(p * gx).sum(axis=1), (p * gy).sum(axis=1)], axis=1), z


def size_regression(W, hs):
    """OETR.size_regression, src/model.py:188-191 / :59-63."""
    h = np.maximum(hs[:, 0, :] @ W["tlbr_reg.0.weight"].T, 0.0)
    o = h @ W["tlbr_reg.2.weight"].T + W["tlbr_reg.2.bias"]
    return 1.0 / (1.0 + np.exp(-o))


def box_tlbr_to_xyxy(cxy, tlbr, max_h, max_w, clamp=True):
    """src/models/utils.py:16-28 (clamp=True, forward_dummy) / src/model.py:193-211 (clamp=False, forward)."""
    t, l, b, r = (tlbr[:, i] for i in range(4))
    x, y = cxy[:, 0], cxy[:, 1]
    box = np.stack([x - l * max_w, y - t * max_h, x + r * max_w, y + b * max_h], axis=1)
    if clamp:
        box[:, 0::2] = np.clip(box[:, 0::2], 0.0, max_w)
        box[:, 1::2] = np.clip(box[:, 1::2], 0.0, max_h)
    return box


# --------------------------------------------------------------------------------------------------------
# the whole hot path
# --------------------------------------------------------------------------------------------------------
def hot_path(weights, feat1, feat2, img_hw1, img_hw2, attention="linear", clamp=True, dtype=np.float64,
             max_shape=(100, 100), return_layers=False, rnd=_identity, mask1=None, mask2=None):
    """feature_correlation + center_estimation + size_regression + box_tlbr_to_xyxy
    (src/model.py:240-250).  feat1 [N,256,hf1,wf1], feat2 [N,256,hf2,wf2] (outputs of input_proj2).
    Returns a dict of every stage boundary."""
    W = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    f1 = np.asarray(feat1, dtype=dtype)
    f2 = np.asarray(feat2, dtype=dtype)
    hf1, wf1 = f1.shape[2:]
    hf2, wf2 = f2.shape[2:]
    pe = pe_table(max_shape, dtype=dtype)
    pos1 = pe[:, :hf1, :wf1]
    pos2 = pe[:, :hf2, :wf2]
    res = query_transformer(W, f1, f2, pos1, pos2, attention, return_layers=return_layers, rnd=rnd,
                            mask0=mask1, mask1=mask2)
    hs1, hs2, mem1, mem2 = res[:4]
    cxy1, z1 = center_estimation(W, hs1, mem1, hf1, wf1, img_hw1[0], mask1)
    cxy2, z2 = center_estimation(W, hs2, mem2, hf2, wf2, img_hw2[0], mask2)
    tlbr1 = size_regression(W, hs1)
    tlbr2 = size_regression(W, hs2)
    out = dict(hs1=hs1[:, 0], hs2=hs2[:, 0], memory1=mem1, memory2=mem2, cxy1=cxy1, cxy2=cxy2,
               tlbr1=tlbr1, tlbr2=tlbr2, logits1=z1, logits2=z2,
               box1=box_tlbr_to_xyxy(cxy1, tlbr1, img_hw1[0], img_hw1[1], clamp),
               box2=box_tlbr_to_xyxy(cxy2, tlbr2, img_hw2[0], img_hw2[1], clamp),
               box1_raw=box_tlbr_to_xyxy(cxy1, tlbr1, img_hw1[0], img_hw1[1], False),
               box2_raw=box_tlbr_to_xyxy(cxy2, tlbr2, img_hw2[0], img_hw2[1], False))
    if return_layers:
        out["layers"] = res[4]
    return out
