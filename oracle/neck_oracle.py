"""CPU oracle for the OETR neck -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement (fp64 by default) of what the reference runs between the ResNet backbone and the hot path:
    feat = input_proj(backbone_out)            src/model.py:45-47,118-119     1x1 conv 1024 -> 256, bias
    feat = patchmerging(feat)                  src/models/backbone.py:53-67   LayerNorm over channels, then three
                                                                               stride-2 convolutions k = 4 / 8 / 16,
                                                                               padding (k-2)/2, 256 -> 256/128/128,
                                                                               concatenated on channels
    feat = input_proj2(feat)                   src/model.py:48-50,123-124     1x1 conv 512 -> 256, bias
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Parity status: pinned against the UNMODIFIED reference modules (tests/golden/make_neck_golden.py ->
tests/golden/neck_*.npz; tests/test_neck_oracle.py).

Weights: dict {state_dict key -> ndarray} with the reference's names (NECK_ORDER in oetr_b200/weights.py).
"""
import numpy as np

LN_EPS = 1e-5
PATCH_SIZES = (4, 8, 16)          # src/model.py:55


def _identity(a):
    return a


def conv1x1(x, weight, bias, rnd=_identity):
    """nn.Conv2d(kernel_size=1): x [N,Cin,H,W], weight [Cout,Cin,1,1] -> [N,Cout,H,W]."""
    w = weight.reshape(weight.shape[0], -1)
    return np.einsum("oc,nchw->nohw", rnd(w), rnd(x)) + bias[None, :, None, None]


def layer_norm_channels(x, gamma, beta, eps=LN_EPS):
    """PatchMerging's norm: rearrange to tokens, nn.LayerNorm(dim) (biased variance), rearrange back
    (src/models/backbone.py:58-60).  x [N,C,H,W]."""
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma[None, :, None, None] + beta[None, :, None, None]


def conv_stride2(x, weight, bias, rnd=_identity):
    """nn.Conv2d(dim, out, kernel_size=ps, stride=2, padding=(ps-2)//2) (src/models/backbone.py:44-51):
    out[n,o,oy,ox] = b[o] + sum_{c,ky,kx} w[o,c,ky,kx] * xpad[n,c,2*oy+ky,2*ox+kx]; output size floor(H/2)."""
    n, c, h, w_ = x.shape
    o, _, k, _ = weight.shape
    p = (k - 2) // 2
    ho = (h + 2 * p - k) // 2 + 1
    wo = (w_ + 2 * p - k) // 2 + 1
    xp = np.zeros((n, c, h + 2 * p, w_ + 2 * p), dtype=x.dtype)
    xp[:, :, p:p + h, p:p + w_] = rnd(x)
    wr = rnd(weight)
    out = np.zeros((n, o, ho, wo), dtype=x.dtype)
    for ky in range(k):                 # tap decomposition: 1x1 products of shifted, strided views
        for kx in range(k):
            v = xp[:, :, ky:ky + 2 * ho - 1:2, kx:kx + 2 * wo - 1:2]
            out += np.einsum("oc,nchw->nohw", wr[:, :, ky, kx], v)
    return out + bias[None, :, None, None]


def patch_merging(W, x, prefix="patchmerging.", rnd=_identity):
    """PatchMerging.forward, src/models/backbone.py:53-67."""
    xn = layer_norm_channels(x, W[prefix + "norm.weight"], W[prefix + "norm.bias"])
    outs = [conv_stride2(xn, W[prefix + "reductions.%d.weight" % i], W[prefix + "reductions.%d.bias" % i], rnd)
            for i in range(len(PATCH_SIZES))]
    return np.concatenate(outs, axis=1)


def neck(W, backbone_out, dtype=np.float64, rnd_proj=_identity, rnd_conv=_identity, rnd_proj2=_identity,
         return_stages=False):
    """backbone_out [N,1024,H,W] -> feat [N,256,H//2,W//2]  (src/model.py:118-124, one image set).
    rnd_*: operand-rounding models of the tensor-core path (identity = the reference's arithmetic)."""
    Wd = {k: np.asarray(v, dtype=dtype) for k, v in W.items()}
    x = np.asarray(backbone_out, dtype=dtype)
    p = conv1x1(x, Wd["input_proj.weight"], Wd["input_proj.bias"], rnd_proj)
    m = patch_merging(Wd, p, rnd=rnd_conv)
    f = conv1x1(m, Wd["input_proj2.weight"], Wd["input_proj2.bias"], rnd_proj2)
    if return_stages:
        return {"proj": p, "merged": m, "feat": f}
    return f
