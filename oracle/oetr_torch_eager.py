"""Eager-PyTorch restatement of the OETR hot path -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

The same algorithm as oracle/oetr_oracle.py (which is pinned to the reference's outputs), written with the
torch operators the reference itself is made of (F.layer_norm, F.elu, einsum, F.gelu, F.conv2d, F.group_norm,
softmax), fp32, TF32 off.  Purpose: the "GPU bar to beat" of SURVEY.md section 8(d) -- what the path costs as
~700 eager launches on the same B200 -- reported by bench.py as `gpu_eager_baseline`; /root/reference does not
exist on the GPU box, so the reference's own modules cannot be timed there.  tests/test_oracle_golden.py checks
this file against the numpy oracle on CPU.  Only tests/ and bench.py import it.

Reference lines: src/models/transformer.py:104-142 (encoder layer), :55-72, :224-255 (decoder),
:313-383 (QueryTransformer); src/models/linear_attention.py:22-50; src/model.py:145-191; src/models/utils.py:16-28.
"""
import torch
import torch.nn.functional as F

from . import oetr_oracle as orc

C, NH, HD = orc.D_MODEL, orc.NHEAD, orc.HEAD_DIM


def _lin_attn(q, k, v):
    """linear_attention.py:22-50.  q [N,L,H,D]; k, v [N,S,H,D]."""
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    s_len = v.size(1)
    v = v / s_len
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + orc.ATTN_EPS)
    return torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * s_len


def _full_attn(q, k, v):
    """linear_attention.py:59-87 (masks None, dropout off): softmax(q k^T / sqrt(D)) v per head."""
    QK = torch.einsum("nlhd,nshd->nlsh", q, k)
    A = torch.softmax(QK / q.size(3) ** 0.5, dim=2)
    return torch.einsum("nlsh,nshd->nlhd", A, v)


def _encoder_layer(W, p, x, src, x_pos, s_pos, attention="linear"):
    n, l, _ = x.shape
    query = F.layer_norm(x, (C,), W[p + "pre_norm_q.weight"], W[p + "pre_norm_q.bias"]) + x_pos
    kv = F.layer_norm(src, (C,), W[p + "pre_norm_kv.weight"], W[p + "pre_norm_kv.bias"]) + s_pos
    q = F.linear(query, W[p + "q_proj.weight"]).view(n, l, NH, HD)
    k = F.linear(kv, W[p + "k_proj.weight"]).view(n, -1, NH, HD)
    v = F.linear(kv, W[p + "v_proj.weight"]).view(n, -1, NH, HD)
    att = _lin_attn(q, k, v) if attention == "linear" else _full_attn(q, k, v)
    x = x + F.linear(att.reshape(n, l, C), W[p + "merge.weight"])
    h = F.gelu(F.linear(F.layer_norm(x, (C,), W[p + "norm2.weight"], W[p + "norm2.bias"]), W[p + "mlp.0.weight"]))
    return x + F.linear(h, W[p + "mlp.2.weight"])


def _mha(W, p, q_in, k_in, v_in):
    n = q_in.size(0)
    q = F.linear(q_in, W[p + "q_proj.weight"], W[p + "q_proj.bias"]).view(n, -1, NH, HD)
    k = F.linear(k_in, W[p + "k_proj.weight"], W[p + "k_proj.bias"]).view(n, -1, NH, HD)
    v = F.linear(v_in, W[p + "v_proj.weight"], W[p + "v_proj.bias"]).view(n, -1, NH, HD)
    return F.linear(_lin_attn(q, k, v).reshape(n, -1, C), W[p + "merge.weight"])


def _decoder_layer(W, p, tgt, memory, qe, m_pos):
    t2 = F.layer_norm(tgt, (C,), W[p + "norm1.weight"], W[p + "norm1.bias"])
    qk = t2 + qe
    tgt = tgt + _mha(W, p + "self_attn.", qk, qk, t2)
    t2 = F.layer_norm(tgt, (C,), W[p + "norm2.weight"], W[p + "norm2.bias"])
    tgt = tgt + _mha(W, p + "multihead_attn.", t2 + qe, memory + m_pos, memory)
    t2 = F.layer_norm(tgt, (C,), W[p + "norm3.weight"], W[p + "norm3.bias"])
    return tgt + F.linear(F.relu(F.linear(t2, W[p + "mlp.0.weight"])), W[p + "mlp.2.weight"])


def _center(W, hs, memory, hf, wf, img_h):
    n = memory.size(0)
    att = torch.einsum("blc,bnc->bln", memory, hs)
    heat = (memory * att).transpose(1, 2).reshape(n, C, hf, wf)
    y = F.conv2d(heat, W["heatmap_conv.0.weight"], W["heatmap_conv.0.bias"], padding=1)
    y = F.relu(F.group_norm(y, 32, W["heatmap_conv.1.weight"], W["heatmap_conv.1.bias"]))
    z = F.conv2d(y, W["heatmap_conv.3.weight"], W["heatmap_conv.3.bias"]).reshape(n, hf * wf)
    p = torch.softmax(z, dim=1)
    stride = img_h // hf
    ys, xs = torch.meshgrid(torch.arange(hf, device=z.device), torch.arange(wf, device=z.device), indexing="ij")
    gx = (xs.reshape(-1).to(z.dtype) + 0.5) * stride
    gy = (ys.reshape(-1).to(z.dtype) + 0.5) * stride
    return torch.stack([(p * gx).sum(dim=1), (p * gy).sum(dim=1)], dim=1)


def _boxes(cxy, tlbr, max_h, max_w, clamp):
    t, l, b, r = tlbr.unbind(dim=1)
    x, y = cxy.unbind(dim=1)
    box = torch.stack([x - l * max_w, y - t * max_h, x + r * max_w, y + b * max_h], dim=1)
    if clamp:
        box = torch.stack([box[:, 0].clamp(0, max_w), box[:, 1].clamp(0, max_h),
                           box[:, 2].clamp(0, max_w), box[:, 3].clamp(0, max_h)], dim=1)
    return box


def prepare(weights, device, dtype=torch.float32, max_shape=(100, 100)):
    """Weights (state-dict-keyed numpy arrays) and the position table as torch tensors on `device`."""
    W = {k: torch.as_tensor(v).to(device=device, dtype=dtype) for k, v in weights.items()}
    W["_pe"] = torch.as_tensor(orc.pe_table(max_shape)).to(device=device, dtype=dtype)
    return W


@torch.no_grad()
def hot_path(W, feat1, feat2, img_hw1, img_hw2, clamp=True, attention="linear"):
    """feat1 [N,256,hf1,wf1], feat2 [N,256,hf2,wf2] tensors on W's device.  Returns (box1, box2) [N,4]."""
    n = feat1.size(0)
    hf1, wf1 = feat1.shape[2:]
    hf2, wf2 = feat2.shape[2:]
    x = [feat1.flatten(2).transpose(1, 2), feat2.flatten(2).transpose(1, 2)]
    pos = [W["_pe"][:, :hf1, :wf1].flatten(1).t(), W["_pe"][:, :hf2, :wf2].flatten(1).t()]
    for i in range(orc.N_ENCODER):
        p = "transformer.encoder.%d." % i
        if i % 2 == 0:
            x = [_encoder_layer(W, p, x[0], x[0], pos[0], pos[0], attention), _encoder_layer(W, p, x[1], x[1], pos[1], pos[1], attention)]
        else:
            x = [_encoder_layer(W, p, x[0], x[1], pos[0], pos[1], attention), _encoder_layer(W, p, x[1], x[0], pos[1], pos[0], attention)]
    out = []
    geo = ((hf1, wf1, img_hw1), (hf2, wf2, img_hw2))
    for k in range(2):
        qe = W["query_embed%d.weight" % (k + 1)][None].expand(n, 1, C)
        t = torch.zeros(n, 1, C, device=feat1.device, dtype=feat1.dtype)
        for j in range(orc.N_DECODER):
            t = _decoder_layer(W, "transformer.decoder.layers.%d." % j, t, x[k], qe, pos[k])
        hf, wf, hw = geo[k]
        cxy = _center(W, t, x[k], hf, wf, hw[0])
        tlbr = torch.sigmoid(F.linear(F.relu(F.linear(t[:, 0], W["tlbr_reg.0.weight"])), W["tlbr_reg.2.weight"],
                                      W["tlbr_reg.2.bias"]))
        out.append(_boxes(cxy, tlbr, hw[0], hw[1], clamp))
    return out[0], out[1]
